"""Training step of the Myriad hot path on the C-ABI kernels: forward with saved activations, backward, data-parallel
gradient all-reduce and fused AdamW (reference: Myriad.forward myriad.py:377-431 under base_task.py:233-271 and DDP
runner_base.py:96-98).

Only the reference's trainable parameters receive gradients (runner_base.py:111-119): `expert_adaptor.conv{1,2}`,
`VEInstructor.meta_net.*`, `VETokenizer.meta_net.*`, `VETokenizer.base_prompts`, and with use_lora the peft LoRA A/B of
q_proj / v_proj. Everything else is frozen: activations get dgrad only, no wgrad is ever computed for frozen weights, and
the ViT runs forward-only (nothing trainable sits below it).

All trainable parameters live in ONE flat fp32 buffer (`flat_params`) with a matching flat gradient buffer: one NCCL
all-reduce per optimizer step over NVLink (untouched slices stay zero — DDP find_unused_parameters semantics), one fused
AdamW launch. Conv filters are stored [Cout, kh, kw, Cin] (NHWC order) inside the flat buffer; `export_state_dict`
permutes back to the reference layout.

Gradient activations are fp16 GEMM operands, so the loss is scaled (GradScaler semantics, runner_base.py:141-149) and
weight-gradient epilogues / the optimizer unscale in fp32.
"""
import math
import os

import torch

from . import kernels as K
from .engine import F16, F32, MyriadEngine, _Obj
from .synthetic import CONV_CHANNELS, CONV_IDX


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class MyriadTrainer(MyriadEngine):
    FUSED_LLAMA = False  # plain weight layout: the backward needs the pre-activation gate/up values and xa = x A^T

    def __init__(self, sd, dims, device="cuda:0", max_batch=8, max_seq=512, loss_scale=1024.0, lr=1e-4, betas=(0.9, 0.999),
                 eps=1e-8, weight_decay=0.05, lora_dropout=0.0, dropout_seed=0):
        # peft's lora_dropout (0.05 in the reference's LoraConfig, myriad.py:171-178): the only stochastic op of the training hot
        # path. 0.0 = off (parity runs against the oracle use a shared mask or no dropout). The mask stream is keyed by
        # (dropout_seed, forward count, layer, branch), see _lora_drop_offset.
        self.lora_dropout, self.dropout_seed, self.fwd_count = float(lora_dropout), int(dropout_seed), 0
        self.overlap_allreduce = True
        self._early = None
        # MYR_LORA_FUSED=0: the LoRA branches as tcgen05 GEMMs + separate dropout launches (16 launches per layer instead of 5)
        self.lora_fused = os.environ.get("MYR_LORA_FUSED", "1") != "0" and dims.lora_r == 8 and dims.llama.hidden % 256 == 0
        self.supervised_rows_only = True  # lm_head / clamp-CE only on the rows that carry a target (False: all positions, as the reference)
        self._sd_for_flat = sd
        super().__init__(sd, dims, device, max_batch, max_seq)
        self.loss_scale = float(loss_scale)
        self.hp = dict(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, wd=weight_decay)
        self.opt_step = 0
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.found_inf = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self._saved = None
        # GradScaler semantics for the internal fp16 backward (runner_base.py:141-149): the scale every weight-gradient epilogue
        # divides out again is driven by the device-side inf / nan flag of the optimizer step. The flag is read back without
        # stalling the launch queue (pinned copy + event, examined one step later); a skipped step is taken back out of
        # opt_step so AdamW's bias correction counts applied updates only, like torch.optim.AdamW under GradScaler.step.
        from .optim import DynamicLossScale
        self.scaler = DynamicLossScale(init_scale=self.loss_scale)
        self._flag_ring = [(torch.zeros(1, dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(4)] if self.dev.type == "cuda" else []
        self._flag_pending = []
        self.skipped_steps = 0

    # ------------------------------------------------------------------------------- flat parameter space
    def _prep_experts(self, sd):
        """Called by MyriadEngine.__init__: builds the flat fp32 buffer first, then the engine's device views/copies."""
        d, dev = self.d, self.dev
        spec = []  # (key, ref_shape, kind)
        spec += [("expert_adaptor.conv1.weight", (d.adaptor_rank, d.vit.dim), "plain"),
                 ("expert_adaptor.conv2.weight", (d.vit.dim, d.adaptor_rank), "plain")]
        for mod, on, ho, hk in (("VEInstructor", d.use_instructor, 768, 1), ("VETokenizer", d.use_tokenizer, 4096, 5)):
            if not on:
                continue
            for j, idx in enumerate(CONV_IDX):
                spec += [("%s.meta_net.%d.weight" % (mod, idx), (CONV_CHANNELS[j + 1], CONV_CHANNELS[j], 3, 3), "conv"),
                         ("%s.meta_net.%d.bias" % (mod, idx), (CONV_CHANNELS[j + 1],), "plain")]
            spec += [("%s.meta_net.15.weight" % mod, (ho, 1024, hk, hk), "conv"), ("%s.meta_net.15.bias" % mod, (ho,), "plain")]
        if d.use_tokenizer:
            spec += [("VETokenizer.base_prompts", (9, 4096), "plain")]
        if d.lora_r:
            for i in range(d.llama.layers):
                p = "llama_model.base_model.model.model.layers.%d.self_attn." % i
                # A of q and v adjacent (one [2r, D] block), then B_q, B_v
                spec += [(p + "q_proj.lora_A.default.weight", (d.lora_r, d.llama.hidden), "plain"),
                         (p + "v_proj.lora_A.default.weight", (d.lora_r, d.llama.hidden), "plain"),
                         (p + "q_proj.lora_B.default.weight", (d.llama.hidden, d.lora_r), "plain"),
                         (p + "v_proj.lora_B.default.weight", (d.llama.hidden, d.lora_r), "plain")]
        self.segments, off = {}, 0
        for key, shape, kind in spec:
            n = _numel(shape)
            self.segments[key] = (off, shape, kind)
            off += (n + 3) // 4 * 4  # keep every segment 16-byte aligned
        self.flat_params = torch.zeros(off, device=dev, dtype=F32)
        self.flat_grads = torch.zeros(off, device=dev, dtype=F32)
        self.wd_mask = torch.zeros(off, device=dev, dtype=torch.uint8)
        for key, (o, shape, kind) in self.segments.items():
            t = sd[key].to(dev, F32)
            if kind == "conv":
                t = t.permute(0, 2, 3, 1)
            self.flat_params[o:o + _numel(shape)].copy_(t.reshape(-1))
            # runner_base.py:115: no weight decay for ndim < 2 or bias / ln / bn parameters
            from .optim import no_weight_decay
            if not no_weight_decay(key, len(shape)):
                self.wd_mask[o:o + _numel(shape)] = 1
        self.instw = self._conv_views("VEInstructor", 1) if d.use_instructor else None
        self.tokw = self._conv_views("VETokenizer", 5) if d.use_tokenizer else None
        if self.tokw is not None:
            self.tokw.base_prompts = self.param("VETokenizer.base_prompts")
        self.refresh_trainables()

    def param(self, key, buf=None):
        o, shape, kind = self.segments[key]
        buf = self.flat_params if buf is None else buf
        v = buf[o:o + _numel(shape)]
        if kind == "conv":
            return v.view(shape[0], shape[2], shape[3], shape[1])
        return v.view(*shape)

    def grad(self, key):
        return self.param(key, self.flat_grads)

    def _conv_views(self, mod, head_k):
        W = _Obj()
        W.mod, W.head_k = mod, head_k
        W.direct, W.gemm = [], []
        for j, idx in enumerate(CONV_IDX):
            wk, bk = "%s.meta_net.%d.weight" % (mod, idx), "%s.meta_net.%d.bias" % (mod, idx)
            cin, cout = CONV_CHANNELS[j], CONV_CHANNELS[j + 1]
            if j < 3:
                W.direct.append((self.param(wk), self.param(bk), cin, cout))
            else:
                W.gemm.append([None, None, cin, cout])  # fp16 copies filled by refresh_trainables
        W.head_w = W.head_b = None
        return W

    def refresh_trainables(self):
        """fp32 master -> fp16 GEMM operands (after load and after every optimizer step)."""
        dev = self.dev

        def h(key, rows):
            src = self.param(key).reshape(rows, -1)
            dst = torch.empty(src.shape, device=dev, dtype=F16)
            K.copy_rows(src, dst, 1, rows, src.shape[1], src.shape[1], 0, src.shape[1], 0)
            return dst

        for W in (self.instw, self.tokw):
            if W is None:
                continue
            for j in (3, 4):
                idx = CONV_IDX[j]
                W.gemm[j - 3][0] = h("%s.meta_net.%d.weight" % (W.mod, idx), CONV_CHANNELS[j + 1])
                W.gemm[j - 3][1] = self._h1("%s.meta_net.%d.bias" % (W.mod, idx))
            W.head_w = h("%s.meta_net.15.weight" % W.mod, self.segments["%s.meta_net.15.weight" % W.mod][1][0])
            W.head_b = self._h1("%s.meta_net.15.bias" % W.mod)
        if hasattr(self, "vitw"):
            self._refresh_late()

    def _h1(self, key):
        src = self.param(key)
        n = src.numel()
        if n % 4:
            return src.to(F16)  # tiny odd-sized vectors only (never hit with the reference channel counts)
        dst = torch.empty(n, device=self.dev, dtype=F16)
        K.copy_rows(src, dst, 1, 1, n, n, 0, n, 0)
        return dst

    def _refresh_late(self):
        self.vitw.ad1 = self.param("expert_adaptor.conv1.weight")
        self.vitw.ad2 = self.param("expert_adaptor.conv2.weight")
        if self.d.lora_r and hasattr(self, "llw"):
            r, D = self.d.lora_r, self.d.llama.hidden
            for i, L in enumerate(self.llw.layers):
                p = "llama_model.base_model.model.model.layers.%d.self_attn." % i
                o = self.segments[p + "q_proj.lora_A.default.weight"][0]
                a32 = self.flat_params[o:o + 2 * r * D].view(2 * r, D)
                L.lora.a = torch.empty(2 * r, D, device=self.dev, dtype=F16)
                K.copy_rows(a32, L.lora.a, 1, 2 * r, D, D, 0, D, 0)
                for nm, attr in (("q_proj", "bq"), ("v_proj", "bv")):
                    b32 = self.param(p + nm + ".lora_B.default.weight")
                    b16 = torch.empty(D, r, device=self.dev, dtype=F16)
                    K.copy_rows(b32.reshape(1, -1), b16.reshape(1, -1), 1, 1, D * r, D * r, 0, D * r, 0)
                    setattr(L.lora, attr, b16)
                L.lora.scale = self.d.lora_alpha / self.d.lora_r

    def _prep_llama(self, sd):
        super()._prep_llama(sd)
        if self.d.lora_r:  # trainer keeps B unscaled (scale applied in the GEMM epilogue), refreshed from the flat buffer
            for L in self.llw.layers:
                L.lora.scale = self.d.lora_alpha / self.d.lora_r
        # The frozen projections are also kept TRANSPOSED ([K_in, F_out] -> rows = input features): the input-gradient GEMM
        # dX = dY W then has both operands K-major and runs on the CTA-pair kernel (csrc/gemm2.cu) like the forward, instead of
        # the MN-major operand path. Costs one more fp16 copy of the LLaMA linear weights (12.9 GB of the 180 GB).
        self.dgrad_transposed = os.environ.get("MYR_DGRAD_T", "1") != "0"
        if self.dgrad_transposed:
            for L in self.llw.layers:
                L.wqkv_t, L.wo_t = L.wqkv.t().contiguous(), L.wo.t().contiguous()
                L.wgu_t, L.wd_t = L.wgu.t().contiguous(), L.wd.t().contiguous()
        self._refresh_late()

    def _dgrad(self, dy, L, name, **kw):
        """dX = dY W for one of the frozen LLaMA projections (name in wqkv / wo / wgu / wd)."""
        if self.dgrad_transposed:
            return K.gemm(dy, getattr(L, name + "_t"), **kw)
        return K.gemm(dy, getattr(L, name), w_mn_major=True, **kw)

    def export_state_dict(self):
        out = {}
        for key, (o, shape, kind) in self.segments.items():
            t = self.param(key)
            out[key] = (t.permute(0, 3, 1, 2) if kind == "conv" else t).contiguous().clone()
        return out

    def export_flat(self, buf):
        """{reference key: tensor in the reference layout} view of any flat buffer shaped like flat_params."""
        out = {}
        for key in self.segments:
            t = self.param(key, buf)
            out[key] = (t.permute(0, 3, 1, 2) if self.segments[key][2] == "conv" else t).contiguous().clone()
        return out

    def import_flat(self, buf, tensors):
        for key, (off, shape, kind) in self.segments.items():
            if key not in tensors:
                continue
            t = tensors[key].to(self.dev, F32)
            if kind == "conv":
                t = t.permute(0, 2, 3, 1)
            buf[off:off + _numel(shape)].copy_(t.reshape(-1))

    def import_state_dict(self, sd):
        """Trainable parameters in the reference layout (a `ckpt:` file's "model" entry) -> flat buffer + fp16 operand copies."""
        self.import_flat(self.flat_params, sd)
        self.refresh_trainables()

    def export_grads(self, unscale=True):
        out = {}
        s = 1.0  # weight-gradient epilogues already divide by loss_scale
        for key, (o, shape, kind) in self.segments.items():
            t = self.grad(key)
            out[key] = (t.permute(0, 3, 1, 2) if kind == "conv" else t).contiguous().clone() * s
        return out

    # ------------------------------------------------------------------------------------ LoRA (trainer form)
    def _llama_layer(self, L, li, h32, bufs, B, S, pos, kv_len, cache_off, cache_off_dev, Skv, causal):
        """Inference launch sequence of MyriadEngine._llama_layer, with the trainer's unscaled LoRA B (alpha / r applied in
        the GEMM epilogue) so evaluation between optimizer steps uses the live parameters."""
        l = self.d.llama
        D, H, dh, T = l.hidden, l.heads, l.head_dim, B * S
        x16, qkv, ctx, gu, act = bufs
        kc, vc = self.kcache[li], self.vcache[li]
        K.norm(h32, L.n1, None, l.eps, rms=True, out16=x16)
        K.gemm(x16, L.wqkv, out=qkv)
        if L.lora is not None:
            r = self.d.lora_r
            xa = K.gemm(x16, L.lora.a)
            K.gemm(xa[:, :r], L.lora.bq, res=qkv[:, :D], out=qkv[:, :D], T=T, K=r, alpha=L.lora.scale)
            K.gemm(xa[:, r:], L.lora.bv, res=qkv[:, 2 * D:], out=qkv[:, 2 * D:], T=T, K=r, alpha=L.lora.scale)
        K.rope_cache(qkv, B, S, H, dh, pos, self.llw.cos, self.llw.sin, kc, vc, cache_off=cache_off, cache_off_dev=cache_off_dev)
        cs = (kc.stride(1), kc.stride(0), dh)
        K.attention(qkv, kc, vc, ctx, B, H, S, Skv, dh, 1.0 / math.sqrt(dh), (3 * D, S * 3 * D, dh), cs, cs, (D, S * D, dh),
                    causal=causal, q_off=0, kv_len=kv_len)
        K.gemm(ctx, L.wo, res=h32, out=h32)
        K.norm(h32, L.n2, None, l.eps, rms=True, out16=x16)
        K.gemm(x16, L.wgu, out=gu)
        K.swiglu(gu, act, T, l.inter)
        K.gemm(act, L.wd, res=h32, out=h32)

    # ------------------------------------------------------------------------------------------ small helpers
    def _e(self, *shape, dtype=F16):
        return torch.empty(*shape, device=self.dev, dtype=dtype)

    def _cast16(self, src32, rows, D, src_ld=None, src_gs=0, groups=1):
        """fp32 rows -> contiguous fp16 [groups * rows, D] (GEMM operand)."""
        dst = self._e(groups * rows, D)
        K.copy_rows(src32, dst, groups, rows, D, D if src_ld is None else src_ld, src_gs, D, rows * D)
        return dst

    def _norm_bwd(self, x, dy, gamma, eps, rms=False, add=None, want16=True):
        out32 = self._e(*x.shape, dtype=F32)
        out16 = self._e(*x.shape) if want16 else None
        K.norm_bwd(x, dy, gamma, eps, rms=rms, add=add, out32=out32, out16=out16)
        return out32, out16

    # --------------------------------------------------------------------------------------- attention backward
    ATTN_BWD_WS_BYTES = int(os.environ.get("MYR_ATTN_BWD_WS_MB", "1024")) << 20
    ATTN_BWD_FUSED = os.environ.get("MYR_ATTN_BWD_FUSED", "1") != "0"

    def _attn_bwd(self, q, k, v, dctx, dq, dk, dv, B, H, Sq, Skv, dh, scale, causal, kv_len):
        """q/k/v/dq/dk/dv: (tensor_view, token_stride, batch_stride) with head stride dh; dctx fp16 [B*Sq, H*dh].
        Short sequences (dh 64 / 128, Skv <= 256: every attention of the training step but the Q-Former's 257-key cross-attention)
        take the fused kernel of csrc/attn_bwd.cu. Otherwise
        P is re-materialised: S = Q K^T (batched tcgen05 GEMM, fp32) -> masked softmax -> dV = P^T dO, dP = dO V^T,
        dS = scale * P * (dP - rowsum(dP * P)), dQ = dS K, dK = dS^T Q.
        Heads are independent, so the S x S work buffers (8 bytes per score) are sized for a GROUP of heads that fits
        ATTN_BWD_WS_BYTES and the group loop walks the heads: 164-token training sequences take one pass (27 MB), the sweep's
        S = 2048 at batch 4 takes 5 passes of 7 heads (0.94 GB each) instead of one 4.3 GB allocation per layer."""
        if self.ATTN_BWD_FUSED and K.attn_bwd_small_supported(Sq, Skv, dh):
            # short sequences (the training step's): one fused launch, scores stay on the SM (csrc/attn_bwd.cu)
            K.attn_bwd_small(q, k, v, dctx, dq, dk, dv, B, H, Sq, Skv, dh, scale, causal, kv_len)
            return
        dev = self.dev
        Sp = (Skv + 63) // 64 * 64
        per_head = B * Sq * Sp * 8
        hc = max(1, min(H, self.ATTN_BWD_WS_BYTES // max(1, per_head)))
        n = B * hc * Sq
        S32 = torch.empty(n, Sp, device=dev, dtype=F32)
        P16 = torch.empty(n, Sp, device=dev, dtype=F16)
        dS16 = torch.empty(n, Sp, device=dev, dtype=F16)
        (qt, q_ts, q_bs), (kt, k_ts, k_bs), (vt, v_ts, v_bs) = q, k, v
        (dqt, dq_ts, dq_bs), (dkt, dk_ts, dk_bs), (dvt, dv_ts, dv_bs) = dq, dk, dv
        HD = H * dh
        for h0 in range(0, H, hc):
            hn = min(hc, H - h0)
            o = h0 * dh  # element offset of the group's first head inside a token row
            qv, kv_, vv, dov = qt[..., o:], kt[..., o:], vt[..., o:], dctx[..., o:]  # views: only the data pointer moves
            dqv, dkv, dvv = dqt[..., o:], dkt[..., o:], dvt[..., o:]
            obs = (hn * Sq * Sp, Sq * Sp)
            K.gemm(qv, kv_, out=S32.reshape(-1), out_dtype=F32, T=Sq, F=Skv, K=dh, ldx=q_ts, ldw=k_ts, ldo=Sp,
                   batch=(B, hn, (q_bs, dh), (k_bs, dh), obs))
            K.softmax_rows(S32, P16, B, hn, Sq, Skv, Sp, scale, causal, kv_len)
            # dP = dO V^T (reuse S32)
            K.gemm(dov, vv, out=S32.reshape(-1), T=Sq, F=Skv, K=dh, ldx=HD, ldw=v_ts, ldo=Sp,
                   batch=(B, hn, (Sq * HD, dh), (v_bs, dh), obs))
            K.softmax_bwd_rows(P16, S32, dS16, B * hn * Sq, Sp, scale)
            # dV[key, d] = sum_q P[q, key] dO[q, d]
            K.gemm(P16, dov, out=dvv, x_mn_major=True, w_mn_major=True, T=Skv, F=dh, K=Sq, ldx=Sp, ldw=HD, ldo=dv_ts, bn_hint=64,
                   batch=(B, hn, obs, (Sq * HD, dh), (dv_bs, dh)))
            # dQ[q, d] = sum_key dS[q, key] K[key, d]
            K.gemm(dS16, kv_, out=dqv, w_mn_major=True, T=Sq, F=dh, K=Skv, ldx=Sp, ldw=k_ts, ldo=dq_ts,
                   batch=(B, hn, obs, (k_bs, dh), (dq_bs, dh)))
            # dK[key, d] = sum_q dS[q, key] Q[q, d]
            K.gemm(dS16, qv, out=dkv, x_mn_major=True, w_mn_major=True, T=Skv, F=dh, K=Sq, ldx=Sp, ldw=q_ts, ldo=dk_ts, bn_hint=64,
                   batch=(B, hn, obs, (q_bs, dh), (dk_bs, dh)))

    # --------------------------------------------------------------------------------------------- conv stacks
    def _conv_trunk_train(self, maps, W):
        """Forward of the 5-layer trunk keeping (input, pre-pool activation[, im2col matrix]) per layer."""
        dev = self.dev
        B, Hc = maps.shape[0], maps.shape[2]
        x, saved = maps, []
        for wn, b, cin, cout in W.direct:
            y = torch.empty(B, Hc, Hc, cout, device=dev, dtype=F16)
            K.conv3x3_relu(x, wn, b, y, B, Hc, Hc, cin, cout)
            xn = torch.empty(B, Hc // 2, Hc // 2, cout, device=dev, dtype=F16)
            K.maxpool2(y, xn, B, Hc, Hc, cout)
            saved.append((x, y, None, Hc, cin, cout))
            x, Hc = xn, Hc // 2
        for wg, b, cin, cout in W.gemm:
            cols = torch.empty(B * Hc * Hc, 9 * cin, device=dev, dtype=F16)
            K.im2col(x, cols, B, Hc, Hc, cin, 3, 3, 1)
            y = K.gemm(cols, wg, bias=b, act=K.ACT_RELU)
            xn = torch.empty(B, Hc // 2, Hc // 2, cout, device=dev, dtype=F16)
            K.maxpool2(y, xn, B, Hc, Hc, cout)
            saved.append((x, y, cols, Hc, cin, cout))
            x, Hc = xn, Hc // 2
        return x, saved

    def _conv_trunk_bwd(self, dx, saved, W):
        """dx: fp16 gradient w.r.t. the trunk output [B,7,7,1024]; writes weight/bias grads into the flat buffer."""
        dev = self.dev
        inv_scale = 1.0 / self.loss_scale
        B = dx.shape[0]
        for j in range(4, -1, -1):
            x, y, cols, Hc, cin, cout = saved[j]
            idx = CONV_IDX[j]
            gw, gb = self.grad("%s.meta_net.%d.weight" % (W.mod, idx)), self.grad("%s.meta_net.%d.bias" % (W.mod, idx))
            dy = torch.empty(B, Hc, Hc, cout, device=dev, dtype=F16)
            K.pool_relu_bwd(y, dx, dy, B, Hc, Hc, cout)
            rows = B * Hc * Hc
            if cols is None and j > 0 and (9 * cin) % 8 == 0 and cout % 8 == 0:
                # 16 -> 64 channels at 56 x 56: the direct weight-gradient kernel needs 5.7 ms here (profiles/r1_train_phases.md);
                # as a [64 x 12544] x [12544 x 144] tensor-core GEMM over the im2col matrix it is tens of microseconds
                c2 = torch.empty(rows, 9 * cin, device=dev, dtype=F16)
                K.im2col(x, c2, B, Hc, Hc, cin, 3, 3, 1)
                dy2 = dy.reshape(rows, cout)
                K.gemm(dy2, c2, out=gw.reshape(cout, 9 * cin), x_mn_major=True, w_mn_major=True, T=cout, F=9 * cin, K=rows, bn_hint=64,
                       alpha=inv_scale)
                K.colsum(dy2, cout, 0, 1, rows, cout, gb, scale=inv_scale)
                dx = torch.empty(B, Hc, Hc, cin, device=dev, dtype=F16)
                K.conv3x3_dgrad(dy, W.direct[j][0], dx, B, Hc, Hc, cin, cout)
            elif cols is None:
                K.conv3x3_wgrad(x, dy, gw, gb, B, Hc, Hc, cin, cout, inv_scale)
                if j > 0:
                    dx = torch.empty(B, Hc, Hc, cin, device=dev, dtype=F16)
                    K.conv3x3_dgrad(dy, W.direct[j][0], dx, B, Hc, Hc, cin, cout)
            else:
                dy2 = dy.reshape(rows, cout)
                K.gemm(dy2, cols, out=gw.reshape(cout, 9 * cin), x_mn_major=True, w_mn_major=True, T=cout, F=9 * cin, K=rows,
                       bn_hint=128 if cout >= 128 else 64, alpha=inv_scale)
                K.colsum(dy2, cout, 0, 1, rows, cout, gb, scale=inv_scale)
                dcols = K.gemm(dy2, W.gemm[j - 3][0], w_mn_major=True, F=9 * cin, K=cout)
                dx = torch.empty(B, Hc, Hc, cin, device=dev, dtype=F16)
                K.col2im(dcols, dx, B, Hc, Hc, cin, 3, 3, 1)

    def _ve_head_bwd(self, W, tp, d_tok16, B):
        """Backward of the conv head (1x1 -> 768 for VEInstructor, 5x5 no-pad -> 4096 for VETokenizer) given the fp16
        gradient of its token rows [B * n_tok, Cout]; continues into the trunk."""
        inv_scale = 1.0 / self.loss_scale
        gw, gb = self.grad("%s.meta_net.15.weight" % W.mod), self.grad("%s.meta_net.15.bias" % W.mod)
        rows, cout = d_tok16.shape
        kdim = W.head_k * W.head_k * 1024
        K.gemm(d_tok16, tp.head_in, out=gw.reshape(cout, kdim), x_mn_major=True, w_mn_major=True, T=cout, F=kdim, K=rows, bn_hint=128,
               alpha=inv_scale)
        K.colsum(d_tok16, cout, 0, 1, rows, cout, gb, scale=inv_scale)
        d_in = K.gemm(d_tok16, W.head_w, w_mn_major=True, F=kdim, K=cout)  # [rows, k*k*1024]
        if W.head_k == 1:
            dx = d_in.reshape(B, 7, 7, 1024)
        else:
            dx = self._e(B, 7, 7, 1024)
            K.col2im(d_in, dx, B, 7, 7, 1024, W.head_k, W.head_k, 0)
        self._conv_trunk_bwd(dx, tp.saved, W)

    # ------------------------------------------------------------------------- encode_img with saved activations
    def encode_img(self, image, maps, stage, out=None, out_batch_stride=None):
        tp = getattr(self, "_tape", None)
        if tp is None:
            return super().encode_img(image, maps, stage, out, out_batch_stride)
        d, dev = self.d, self.dev
        B, N, Dv = image.shape[0], d.vit.tokens, d.vit.dim
        Hq, Dl = d.qf.hidden, d.llama.hidden
        W = self.vitw
        tp.vit_x = self.vit_forward(image)  # frozen: forward only (nothing trainable below it)
        enc16 = self._e(B * N, Dv)
        tp.ad_pre = self._e(B * N, Dv, dtype=F32)
        K.norm(tp.vit_x, W.ln_vision[0], W.ln_vision[1], 1e-5, out16=enc16, w1=W.ad1, w2=W.ad2, pre32=tp.ad_pre)
        nq0 = d.qf.num_query
        Q = nq0 + (49 if stage in (1, 2) else 0)
        tp.B, tp.Q, tp.stage = B, Q, stage
        q32 = self._e(B * Q, Hq, dtype=F32)
        K.copy_rows(self.qfw.query_tokens, q32, B, nq0, Hq, Hq, 0, Hq, Q * Hq)
        if stage in (1, 2):
            tp.inst = _Obj()
            trunk, tp.inst.saved = self._conv_trunk_train(maps, self.instw)
            tp.inst.head_in = trunk.reshape(B * 49, 1024)
            K.gemm(tp.inst.head_in, self.instw.head_w, bias=self.instw.head_b, out=q32[nq0:], out_group_rows=49,
                   out_group_stride=Q * Hq, T=B * 49, ldo=Hq)
        h16 = self._qformer_train(q32, enc16, B, Q, tp)
        n_tok = self.num_image_tokens(stage)
        if out is None:
            out = torch.empty(B, n_tok, Dl, device=dev, dtype=F32)
            out_batch_stride = n_tok * Dl
        flat = out.reshape(-1)
        K.gemm(h16, self.qfw.proj_w, bias=self.qfw.proj_b, out=flat, out_group_rows=Q, out_group_stride=out_batch_stride,
               T=B * Q, ldo=Dl)
        if stage in (0, 1):
            Wt = self.tokw
            tp.tok = _Obj()
            trunk, tp.tok.saved = self._conv_trunk_train(maps, Wt)
            tp.tok.head_in = self._e(B * 9, 25 * 1024)
            K.im2col(trunk, tp.tok.head_in, B, 7, 7, 1024, 5, 5, 0)
            K.copy_rows(Wt.base_prompts, flat[Q * Dl:], B, 9, Dl, Dl, 0, Dl, out_batch_stride)
            K.gemm(tp.tok.head_in, Wt.head_w, bias=Wt.head_b, out=flat[(Q + 9) * Dl:], out_group_rows=9,
                   out_group_stride=out_batch_stride, T=B * 9, ldo=Dl)
        return out

    def _qformer_train(self, q32, enc16, B, Q, tp):
        """qformer_forward keeping what the input-gradient pass needs. Every Q-Former weight is frozen, so that is only:
        pre-LayerNorm sums (fp32), q/k/v of both attentions, and the GELU pre-activation."""
        q, W = self.d.qf, self.qfw
        Hd, H = q.hidden, q.heads
        dh = Hd // H
        N = self.d.vit.tokens
        T = B * Q
        h32, h16 = self._e(T, Hd, dtype=F32), self._e(T, Hd)
        ctx, ff = self._e(T, Hd), self._e(T, q.inter)
        tp.q_in = q32
        K.norm(q32, W.emb_ln[0], W.emb_ln[1], q.ln_eps, out16=h16, out32=h32)
        tp.ckv = K.gemm(enc16, W.ckv_w, bias=W.ckv_b)
        ldkv = tp.ckv.shape[1]
        scale = 1.0 / math.sqrt(dh)
        tp.qf = []
        so = (Hd, Q * Hd, dh)
        for L in W.layers:
            S = _Obj()
            S.qkv = K.gemm(h16, L.wqkv, bias=L.bqkv)
            s = (3 * Hd, Q * 3 * Hd, dh)
            K.attention(S.qkv, S.qkv[:, Hd:], S.qkv[:, 2 * Hd:], ctx, B, H, Q, Q, dh, scale, s, s, s, so)
            S.tmp_a = K.gemm(ctx, L.wo, bias=L.bo, res=h32, out_dtype=F32)
            K.norm(S.tmp_a, L.ln_a[0], L.ln_a[1], q.ln_eps, out16=h16, out32=h32)
            if L.cross:
                S.cq = K.gemm(h16, L.cq_w, bias=L.cq_b)
                kk = tp.ckv[:, L.ckv_index * 2 * Hd:]
                ks = (ldkv, N * ldkv, dh)
                K.attention(S.cq, kk, kk[:, Hd:], ctx, B, H, Q, N, dh, scale, so, ks, ks, so)
                S.tmp_c = K.gemm(ctx, L.co_w, bias=L.co_b, res=h32, out_dtype=F32)
                K.norm(S.tmp_c, L.ln_c[0], L.ln_c[1], q.ln_eps, out16=h16, out32=h32)
            S.ff_pre = K.gemm(h16, L.fi_w, bias=L.fi_b)
            K.gelu_fwd(S.ff_pre, ff)
            S.tmp_f = K.gemm(ff, L.fo_w, bias=L.fo_b, res=h32, out_dtype=F32)
            K.norm(S.tmp_f, L.ln_f[0], L.ln_f[1], q.ln_eps, out16=h16, out32=h32)
            tp.qf.append(S)
        return h16

    def _qformer_bwd(self, dh, tp):
        """dh: fp32 [B*Q, hidden] gradient of the Q-Former output. Returns (dq32 [B*Q, hidden] gradient of the query
        embeddings, d_ckv fp16 [B*N, n_cross*2*hidden] gradient of the fused cross-attention K/V projections)."""
        q, W = self.d.qf, self.qfw
        B, Q = tp.B, tp.Q
        Hd, H = q.hidden, q.heads
        dhd = Hd // H
        N = self.d.vit.tokens
        T = B * Q
        scale = 1.0 / math.sqrt(dhd)
        ldkv = tp.ckv.shape[1]
        d_ckv = self._e(B * N, ldkv)
        for L, S in zip(reversed(W.layers), reversed(tp.qf)):
            d32, d16 = self._norm_bwd(S.tmp_f, dh, L.ln_f[0], q.ln_eps)
            d_ff = K.gemm(d16, L.fo_w, w_mn_major=True)
            K.gelu_bwd(S.ff_pre, d_ff, d_ff)
            dh = K.gemm(d_ff, L.fi_w, w_mn_major=True, res=d32, out_dtype=F32)
            if L.cross:
                d32, d16 = self._norm_bwd(S.tmp_c, dh, L.ln_c[0], q.ln_eps)
                d_ctx = K.gemm(d16, L.co_w, w_mn_major=True)
                d_cq = self._e(T, Hd)
                kk = tp.ckv[:, L.ckv_index * 2 * Hd:]
                dkk = d_ckv[:, L.ckv_index * 2 * Hd:]
                self._attn_bwd((S.cq, Hd, Q * Hd), (kk, ldkv, N * ldkv), (kk[:, Hd:], ldkv, N * ldkv), d_ctx,
                               (d_cq, Hd, Q * Hd), (dkk, ldkv, N * ldkv), (dkk[:, Hd:], ldkv, N * ldkv), B, H, Q, N, dhd, scale,
                               False, None)
                dh = K.gemm(d_cq, L.cq_w, w_mn_major=True, res=d32, out_dtype=F32)
            d32, d16 = self._norm_bwd(S.tmp_a, dh, L.ln_a[0], q.ln_eps)
            d_ctx = K.gemm(d16, L.wo, w_mn_major=True)
            dqkv = self._e(T, 3 * Hd)
            st = (3 * Hd, Q * 3 * Hd)
            self._attn_bwd((S.qkv, *st), (S.qkv[:, Hd:], *st), (S.qkv[:, 2 * Hd:], *st), d_ctx, (dqkv, *st), (dqkv[:, Hd:], *st),
                           (dqkv[:, 2 * Hd:], *st), B, H, Q, Q, dhd, scale, False, None)
            dh = K.gemm(dqkv, L.wqkv, w_mn_major=True, res=d32, out_dtype=F32)
        dq32, _ = self._norm_bwd(tp.q_in, dh, W.emb_ln[0], q.ln_eps, want16=False)
        return dq32, d_ckv

    # ------------------------------------------------------------------------------ LLaMA with saved activations
    def _llama_train_fwd(self, embeds32, kv_len, tp, sup_idx=None):
        """-> fp32 logits of the rows named by sup_idx (int32 [R], device) or of all T rows (sup_idx None)."""
        l, dev = self.d.llama, self.dev
        B, S, D = embeds32.shape
        H, dh, T = l.heads, l.head_dim, B * S
        self._ensure_cache(B, S)
        tp.pos = torch.arange(S, device=dev, dtype=torch.int32).repeat(B)
        tp.kv_len, tp.S = kv_len, S
        h = embeds32.reshape(T, D)
        ctx, act = self._e(T, D), self._e(T, l.inter)
        x2 = self._e(T, D)
        tp.ll = []
        for li, L in enumerate(self.llw.layers):
            Sv = _Obj()
            kc, vc = self.kcache[li], self.vcache[li]
            Sv.h_in = h
            Sv.x1 = self._e(T, D)
            K.norm(h, L.n1, None, l.eps, rms=True, out16=Sv.x1)
            Sv.qkv = K.gemm(Sv.x1, L.wqkv)
            Sv.xa = Sv.xd = None
            if L.lora is not None and self.lora_fused:
                # both LoRA branches in two CUDA-core launches (rank 8 is no GEMM shape): xa = drop(x1) A^T, qkv += s * xa B^T; the
                # dropout masks are regenerated wherever the dropped input is needed, so no dropped copy of x1 is kept
                Sv.xa = self._e(T, 2 * self.d.lora_r, dtype=F32)
                K.lora_fwd(Sv.x1, L.lora.a, L.lora.bq, L.lora.bv, Sv.xa, Sv.qkv, 0, 2 * D, L.lora.scale, self.lora_dropout, self.dropout_seed,
                           self._lora_drop_offset(li, 0, T * D), self._lora_drop_offset(li, 1, T * D))
            elif L.lora is not None:
                r = self.d.lora_r
                if self.lora_dropout > 0.0:
                    # peft: each LoRA module drops its own copy of the input: B_q(A_q(drop_q(x))) and B_v(A_v(drop_v(x)))
                    Sv.xa = self._e(T, 2 * r)
                    Sv.xd = []
                    for j in range(2):
                        xd = K.dropout_fwd(Sv.x1, self._e(T, D), self.lora_dropout, self.dropout_seed, self._lora_drop_offset(li, j, T * D))
                        K.gemm(xd, L.lora.a[j * r:(j + 1) * r], out=Sv.xa[:, j * r:], T=T, F=r, K=D, ldo=2 * r)
                        Sv.xd.append(xd)
                else:
                    Sv.xa = K.gemm(Sv.x1, L.lora.a)
                K.gemm(Sv.xa[:, :r], L.lora.bq, res=Sv.qkv[:, :D], out=Sv.qkv[:, :D], T=T, K=r, alpha=L.lora.scale)
                K.gemm(Sv.xa[:, r:], L.lora.bv, res=Sv.qkv[:, 2 * D:], out=Sv.qkv[:, 2 * D:], T=T, K=r, alpha=L.lora.scale)
            K.rope_cache(Sv.qkv, B, S, H, dh, tp.pos, self.llw.cos, self.llw.sin, kc, vc)
            cs = (kc.stride(1), kc.stride(0), dh)
            K.attention(Sv.qkv, kc, vc, ctx, B, H, S, S, dh, 1.0 / math.sqrt(dh), (3 * D, S * 3 * D, dh), cs, cs, (D, S * D, dh),
                        causal=True, q_off=0, kv_len=kv_len)
            Sv.h_mid = K.gemm(ctx, L.wo, res=h, out_dtype=F32)
            K.norm(Sv.h_mid, L.n2, None, l.eps, rms=True, out16=x2)
            Sv.gu = K.gemm(x2, L.wgu)
            K.swiglu(Sv.gu, act, T, l.inter)
            h = K.gemm(act, L.wd, res=Sv.h_mid, out_dtype=F32)
            tp.ll.append(Sv)
        tp.h_final = h
        x16 = self._e(T, D)
        K.norm(h, self.llw.norm, None, l.eps, rms=True, out16=x16)
        tp.sup_idx = sup_idx
        if sup_idx is not None:
            # the reference runs lm_head on every position (modeling_llama.py:690) and the loss then ignores all but the
            # supervised ones (labels -100): only those rows' logits exist here ([R, V] instead of [T, V], R ~ T / 10)
            xs = self._e(sup_idx.numel(), D)
            K.index_rows(x16, xs, sup_idx, D)
            x16 = xs
        return K.gemm(x16, self.llw.lm_head, out_dtype=F32)  # [R | T, V] fp32

    def _lora_drop_offset(self, layer, branch, n):
        """Start of the mask stream of (this forward, layer, branch q / v): streams never overlap within a run."""
        return ((self.fwd_count * self.d.llama.layers + layer) * 2 + branch) * n

    def _llama_train_bwd(self, dlogits16, tp, B):
        """dlogits16 fp16 [T, V] (already multiplied by loss_scale) -> fp32 [T, D] gradient of inputs_embeds; LoRA A/B
        gradients are written (unscaled) into the flat gradient buffer."""
        l = self.d.llama
        S, D, H, dh = tp.S, l.hidden, l.heads, l.head_dim
        T = B * S
        inv_scale = 1.0 / self.loss_scale
        d_x = K.gemm(dlogits16, self.llw.lm_head, w_mn_major=True, out_dtype=F32)
        if tp.sup_idx is not None:  # rows without a target get no gradient from the loss
            d_all = torch.zeros(T, D, device=self.dev, dtype=F32)
            K.index_rows(d_x, d_all, tp.sup_idx, D, scatter=True)
            d_x = d_all
        dh32, dh16 = self._norm_bwd(tp.h_final, d_x, self.llw.norm, l.eps, rms=True)
        att_scale = 1.0 / math.sqrt(dh)
        for li in range(l.layers - 1, -1, -1):
            L, Sv = self.llw.layers[li], tp.ll[li]
            kc, vc = self.kcache[li], self.vcache[li]
            d_act = self._dgrad(dh16, L, "wd")
            d_gu = self._e(T, 2 * l.inter)
            K.swiglu_bwd(Sv.gu, d_act, d_gu, T, l.inter)
            d_x2 = self._dgrad(d_gu, L, "wgu", out_dtype=F32)
            dmid32, dmid16 = self._norm_bwd(Sv.h_mid, d_x2, L.n2, l.eps, rms=True, add=dh32)
            d_ctx = self._dgrad(dmid16, L, "wo")
            dqkv = self._e(T, 3 * D)
            st = (3 * D, S * 3 * D)
            cst = (kc.stride(1), kc.stride(0))
            self._attn_bwd((Sv.qkv, *st), (kc, *cst), (vc, *cst), d_ctx, (dqkv, *st), (dqkv[:, D:], *st), (dqkv[:, 2 * D:], *st),
                           B, H, S, S, dh, att_scale, True, tp.kv_len)
            K.rope_bwd(dqkv, T, H, dh, tp.pos, self.llw.cos, self.llw.sin)
            d_x1 = self._dgrad(dqkv, L, "wqkv", out_dtype=F32)
            if L.lora is not None and self.lora_fused:
                r = self.d.lora_r
                p = "llama_model.base_model.model.model.layers.%d.self_attn." % li
                o = self.segments[p + "q_proj.lora_A.default.weight"][0]
                K.lora_bwd(dqkv, 0, 2 * D, Sv.xa, L.lora.a, L.lora.bq, L.lora.bv, Sv.x1, self._e(T, 2 * r, dtype=F32),
                           self.grad(p + "q_proj.lora_B.default.weight"), self.grad(p + "v_proj.lora_B.default.weight"),
                           self.flat_grads[o:o + 2 * r * D].view(2 * r, D), d_x1, L.lora.scale, inv_scale, self.lora_dropout, self.dropout_seed,
                           self._lora_drop_offset(li, 0, T * D), self._lora_drop_offset(li, 1, T * D))
            elif L.lora is not None:
                r = self.d.lora_r
                p = "llama_model.base_model.model.model.layers.%d.self_attn." % li
                s = L.lora.scale
                d_xa = self._e(T, 2 * r)
                for j, (nm, b16, col) in enumerate((("q_proj", L.lora.bq, 0), ("v_proj", L.lora.bv, 2 * D))):
                    dy = dqkv[:, col:col + D]
                    # dB[d, r] = s * sum_t dy[t, d] * xa[t, r]
                    K.gemm(dy, Sv.xa[:, j * r:], out=self.grad(p + nm + ".lora_B.default.weight"), x_mn_major=True, w_mn_major=True,
                           T=D, F=r, K=T, ldx=3 * D, ldw=2 * r, bn_hint=128, alpha=s * inv_scale)
                    # d_xa[t, r] = s * sum_d dy[t, d] * B[d, r]
                    K.gemm(dy, b16, out=d_xa[:, j * r:], w_mn_major=True, T=T, F=r, K=D, ldx=3 * D, ldo=2 * r, alpha=s)
                # dA[2r, D] = sum_t d_xa[t, :]^T x1[t, :]   (A_q and A_v are adjacent in the flat buffer)
                o = self.segments[p + "q_proj.lora_A.default.weight"][0]
                if Sv.xd is None:
                    K.gemm(d_xa, Sv.x1, out=self.flat_grads[o:o + 2 * r * D].view(2 * r, D), x_mn_major=True, w_mn_major=True, T=2 * r,
                           F=D, K=T, bn_hint=64, alpha=inv_scale)
                    K.gemm(d_xa, L.lora.a, w_mn_major=True, T=T, F=D, K=2 * r, res=d_x1, out=d_x1)
                else:  # with dropout: A sees the dropped inputs, and the input gradient passes through the same mask
                    g_in = self._e(T, D, dtype=F32)
                    for j in range(2):
                        K.gemm(d_xa[:, j * r:], Sv.xd[j], out=self.flat_grads[o + j * r * D:o + (j + 1) * r * D].view(r, D), x_mn_major=True,
                               w_mn_major=True, T=r, F=D, K=T, ldx=2 * r, bn_hint=64, alpha=inv_scale)
                        K.gemm(d_xa[:, j * r:], L.lora.a[j * r:(j + 1) * r], out=g_in, w_mn_major=True, T=T, F=D, K=r, ldx=2 * r)
                        K.dropout_bwd_add(g_in, d_x1, self.lora_dropout, self.dropout_seed, self._lora_drop_offset(li, j, T * D))
            dh32, dh16 = self._norm_bwd(Sv.h_in, d_x1, L.n1, l.eps, rms=True, add=dmid32, want16=li > 0)
        return dh32

    # -------------------------------------------------------------------------------------------- training step
    def forward_backward(self, image, maps, stage, ids_before, ids_after, text_ids, text_mask):
        """Myriad.forward myriad.py:377-431 (host RNG choices made by the caller: `stage`, and which maps) + backward.
        image fp32 [B,3,224,224], maps fp32 [B,1,224,224] (device); ids_* int64 prompt halves; text_ids int64 [B, Lt]
        right-padded with eos, text_mask [B, Lt] 0/1 (CPU). Returns the loss (fp32 0-d device tensor); gradients of every
        trainable parameter are left, unscaled, in `flat_grads` (parameters untouched by this stage stay zero)."""
        d, l, dev = self.d, self.d.llama, self.dev
        B, D = image.shape[0], l.hidden
        inv_scale = 1.0 / self.loss_scale
        K.memset_zero(self.flat_grads)
        self.fwd_count += 1
        tp = self._tape = _Obj()
        try:
            emb = self.build_inputs_embeds(image, maps, stage, ids_before, ids_after, with_bos=True, text_ids=text_ids)
        finally:
            self._tape = None
        Lt = text_ids.shape[1]
        Ltot = emb.shape[1]
        Lw = Ltot - Lt  # bos + wrapped image prompt
        kv_len = (Lw + text_mask.sum(-1)).to(torch.int32).to(dev)
        # targets myriad.py:406-416, shifted by one for next-token prediction (modeling_llama.py:697-703)
        targets = torch.cat([torch.full((B, Lw), -100, dtype=torch.long), text_ids.masked_fill(text_ids == l.eos, -100)], 1)
        shifted = torch.cat([targets[:, 1:], torch.full((B, 1), -100, dtype=torch.long)], 1).reshape(-1)
        sup = torch.nonzero(shifted != -100).reshape(-1)
        if self.supervised_rows_only and 0 < sup.numel():
            sup_idx = sup.to(torch.int32).to(dev)
            shifted = shifted[sup]
        else:
            sup_idx = None
        shifted = shifted.to(dev)
        logits = self._llama_train_fwd(emb, kv_len, tp, sup_idx)
        T = logits.shape[0]
        row_loss, stats, loss_out = self._e(T, dtype=F32), self._e(T, 2, dtype=F32), self._e(2, dtype=F32)
        K.clamp_ce_fwd(logits, shifted, row_loss, stats, loss_out)
        dlogits = self._e(T, l.vocab)
        K.clamp_ce_bwd(logits, shifted, stats, loss_out, self.loss_scale, dlogits)
        d_emb = self._llama_train_bwd(dlogits, tp, B)  # fp32 [T, D]
        if stage == 2:
            self._early_allreduce()  # no VETokenizer in this stage: its (zero) gradients and the LoRA ones are already final
        # ---- inputs_embeds -> image-token groups (myriad.py:249-266 concatenation order: Q-Former tokens, then VETokenizer)
        n_before = 1 + ids_before.shape[-1]
        flat = d_emb.reshape(-1)
        base = n_before * D
        Q = tp.Q
        d_proj16 = self._cast16(flat[base:], Q, D, src_ld=D, src_gs=Ltot * D, groups=B)
        if stage in (0, 1):
            off = base + Q * D
            K.colsum(flat[off:], D, Ltot * D, B, 1, 9 * D, self.grad("VETokenizer.base_prompts").reshape(-1), scale=inv_scale)
            d_tok16 = self._cast16(flat[off + 9 * D:], 9, D, src_ld=D, src_gs=Ltot * D, groups=B)
            self._ve_head_bwd(self.tokw, tp.tok, d_tok16, B)
            self._early_allreduce()
        d_h = K.gemm(d_proj16, self.qfw.proj_w, w_mn_major=True, out_dtype=F32)  # llama_proj is frozen: dX only
        dq32, d_ckv = self._qformer_bwd(d_h, tp)
        if stage in (1, 2):
            Hq, nq0 = d.qf.hidden, d.qf.num_query
            d_inst16 = self._cast16(dq32.reshape(-1)[nq0 * Hq:], 49, Hq, src_ld=Hq, src_gs=Q * Hq, groups=B)
            self._ve_head_bwd(self.instw, tp.inst, d_inst16, B)
        # cross-attention K/V -> ln_vision -> LoraAdaptorV2 weights (the ViT below is frozen: no dX)
        d_enc = K.gemm(d_ckv, self.qfw.ckv_w, w_mn_major=True, out_dtype=F32)
        d_pre, _ = self._norm_bwd(tp.ad_pre, d_enc, self.vitw.ln_vision[0], 1e-5, want16=False)
        rows, Dv, rk = d_pre.shape[0], d.vit.dim, d.adaptor_rank
        scratch = self._e(rows, 2 * rk, dtype=F32)
        K.adaptor_bwd(tp.vit_x, d_pre, self.vitw.ad1, self.vitw.ad2, scratch, self.grad("expert_adaptor.conv1.weight"),
                      self.grad("expert_adaptor.conv2.weight"), rows, Dv, rk, inv_scale)
        return loss_out[0]

    def _early_allreduce(self):
        """Data parallel: the flat gradient buffer is laid out [adaptor | VEInstructor | VETokenizer | LoRA]; from the VETokenizer
        segment on (97 % of the bytes: the 5x5 head's 420 MB) every gradient is final once the tokenizer head's backward has
        run, while the Q-Former / VEInstructor / adaptor backward is still ahead. That tail is all-reduced NOW, asynchronously
        on NCCL's stream, and overlaps the rest of the backward; optimizer_step reduces the small head of the buffer."""
        import torch.distributed as dist
        self._early = None
        if not (self.overlap_allreduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        key = "VETokenizer.meta_net.0.weight"
        if key not in self.segments:
            return
        split = self.segments[key][0]
        work = dist.all_reduce(self.flat_grads[split:], op=dist.ReduceOp.SUM, async_op=True)
        self._early = (split, work)

    def optimizer_step(self, lr=None):
        """DDP gradient averaging (runner_base.py:96-98: all-reduce of the flat buffer over NCCL/NVLink, the bulk of it
        overlapped with the backward, see _early_allreduce) + fused AdamW (runner_base.py:105-139) + refresh of the fp16
        operand copies."""
        from .dp import allreduce_flat_grads
        early, self._early = getattr(self, "_early", None), None
        if early is not None:
            split, work = early
            mean_scale = allreduce_flat_grads(self.flat_grads[:split])
            work.wait()
        else:
            mean_scale = allreduce_flat_grads(self.flat_grads)  # sum over ranks; 1 / world folded into the optimizer's unscale
        self._poll_overflow_flags(block=len(self._flag_pending) >= len(self._flag_ring))
        self.opt_step += 1
        hp = self.hp
        K.adamw_step(self.flat_params, self.flat_grads, self.exp_avg, self.exp_avg_sq, self.wd_mask, hp["lr"] if lr is None else lr,
                     hp["beta1"], hp["beta2"], hp["eps"], hp["wd"], self.opt_step, inv_scale=mean_scale, found_inf=self.found_inf)
        if self._flag_ring:
            host, ev = self._flag_ring[(self.opt_step + self.skipped_steps) % len(self._flag_ring)]
            host.copy_(self.found_inf, non_blocking=True)
            ev.record()
            self._flag_pending.append((host, ev))
        self.refresh_trainables()

    def _poll_overflow_flags(self, block=False):
        """Fold finished steps' inf / nan flags into the loss scale (x0.5 and the step un-counted on overflow, x2 after 2000
        clean steps). `block` waits for the oldest outstanding flag (used when the small ring is full, and by sync_loss_scale)."""
        while self._flag_pending:
            host, ev = self._flag_pending[0]
            if not ev.query():
                if not block:
                    break
                ev.synchronize()
            self._flag_pending.pop(0)
            bad = bool(int(host[0]))
            if bad:
                self.opt_step -= 1
                self.skipped_steps += 1
            self.loss_scale = float(self.scaler.update(bad))
            block = False

    def sync_loss_scale(self):
        """Wait for every outstanding overflow flag (before a checkpoint is written / when a test inspects the scale)."""
        while self._flag_pending:
            self._poll_overflow_flags(block=True)
        return self.loss_scale

    def train_step(self, image, maps, stage, ids_before, ids_after, text_ids, text_mask, lr=None):
        loss = self.forward_backward(image, maps, stage, ids_before, ids_after, text_ids, text_mask)
        self.optimizer_step(lr)
        return loss
