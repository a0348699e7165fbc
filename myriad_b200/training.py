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
    def __init__(self, sd, dims, device="cuda:0", max_batch=8, max_seq=512, loss_scale=1024.0, lr=1e-4, betas=(0.9, 0.999),
                 eps=1e-8, weight_decay=0.05):
        self._sd_for_flat = sd
        super().__init__(sd, dims, device, max_batch, max_seq)
        self.loss_scale = float(loss_scale)
        self.hp = dict(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, wd=weight_decay)
        self.opt_step = 0
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.found_inf = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self._saved = None

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
            if len(shape) >= 2 and "bias" not in key:
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
        if self.d.lora_r:
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
        self._refresh_late()

    def export_state_dict(self):
        out = {}
        for key, (o, shape, kind) in self.segments.items():
            t = self.param(key)
            out[key] = (t.permute(0, 3, 1, 2) if kind == "conv" else t).contiguous().clone()
        return out

    def export_grads(self, unscale=True):
        out = {}
        s = 1.0  # weight-gradient epilogues already divide by loss_scale
        for key, (o, shape, kind) in self.segments.items():
            t = self.grad(key)
            out[key] = (t.permute(0, 3, 1, 2) if kind == "conv" else t).contiguous().clone() * s
        return out

    # ------------------------------------------------------------------------------------ LoRA (trainer form)
    def _llama_layer(self, L, li, h32, bufs, B, S, pos, kv_len, cache_off, cache_off_dev, Skv, causal, save=None):
        """Same launch sequence as MyriadEngine._llama_layer with B unscaled + epilogue alpha; optionally keeps the
        activations the backward needs (buffers are per-layer when `save` is given)."""
        l = self.d.llama
        D, H, dh, T = l.hidden, l.heads, l.head_dim, B * S
        x16, qkv, ctx, gu, act = bufs
        kc, vc = self.kcache[li], self.vcache[li]
        if save is not None:
            save.h_in = h32.clone() if save.clone_h else h32
        K.norm(h32, L.n1, None, l.eps, rms=True, out16=x16)
        K.gemm(x16, L.wqkv, out=qkv)
        xa = None
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
        if save is not None:
            save.h_mid = h32.clone()
            save.x1, save.qkv, save.xa = x16, qkv, xa
        K.norm(h32, L.n2, None, l.eps, rms=True, out16=x16 if save is None else ctx)  # ctx is free again: reuse as x2
        K.gemm(x16 if save is None else ctx, L.wgu, out=gu)
        K.swiglu(gu, act, T, l.inter)
        K.gemm(act, L.wd, res=h32, out=h32)
        if save is not None:
            save.gu = gu

    # --------------------------------------------------------------------------------------- attention backward
    def _attn_bwd(self, q, k, v, dctx, dq, dk, dv, B, H, Sq, Skv, dh, scale, causal, kv_len):
        """q/k/v/dq/dk/dv: (tensor_view, token_stride, batch_stride) with head stride dh; dctx fp16 [B*Sq, H*dh].
        P is re-materialised: S = Q K^T (batched tcgen05 GEMM, fp32) -> masked softmax -> dV = P^T dO, dP = dO V^T,
        dS = scale * P * (dP - rowsum(dP * P)), dQ = dS K, dK = dS^T Q."""
        dev = self.dev
        Sp = (Skv + 63) // 64 * 64
        n = B * H * Sq
        S32 = torch.empty(n, Sp, device=dev, dtype=F32)
        P16 = torch.empty(n, Sp, device=dev, dtype=F16)
        dS16 = torch.empty(n, Sp, device=dev, dtype=F16)
        obs = (H * Sq * Sp, Sq * Sp)
        (qt, q_ts, q_bs), (kt, k_ts, k_bs), (vt, v_ts, v_bs) = q, k, v
        HD = H * dh
        K.gemm(qt, kt, out=S32.reshape(-1), out_dtype=F32, T=Sq, F=Skv, K=dh, ldx=q_ts, ldw=k_ts, ldo=Sp,
               batch=(B, H, (q_bs, dh), (k_bs, dh), obs))
        K.softmax_rows(S32, P16, B, H, Sq, Skv, Sp, scale, causal, kv_len)
        # dP = dO V^T (reuse S32)
        K.gemm(dctx, vt, out=S32.reshape(-1), T=Sq, F=Skv, K=dh, ldx=HD, ldw=v_ts, ldo=Sp,
               batch=(B, H, (Sq * HD, dh), (v_bs, dh), obs))
        K.softmax_bwd_rows(P16, S32, dS16, n, Sp, scale)
        (dqt, dq_ts, dq_bs), (dkt, dk_ts, dk_bs), (dvt, dv_ts, dv_bs) = dq, dk, dv
        # dV[key, d] = sum_q P[q, key] dO[q, d]
        K.gemm(P16, dctx, out=dvt.reshape(-1) if dvt.dim() == 1 else dvt, x_mn_major=True, w_mn_major=True, T=Skv, F=dh, K=Sq,
               ldx=Sp, ldw=HD, ldo=dv_ts, bn_hint=64, batch=(B, H, obs, (Sq * HD, dh), (dv_bs, dh)))
        # dQ[q, d] = sum_key dS[q, key] K[key, d]
        K.gemm(dS16, kt, out=dqt, w_mn_major=True, T=Sq, F=dh, K=Skv, ldx=Sp, ldw=k_ts, ldo=dq_ts,
               batch=(B, H, obs, (k_bs, dh), (dq_bs, dh)))
        # dK[key, d] = sum_q dS[q, key] Q[q, d]
        K.gemm(dS16, qt, out=dkt, x_mn_major=True, w_mn_major=True, T=Skv, F=dh, K=Sq, ldx=Sp, ldw=q_ts, ldo=dk_ts, bn_hint=64,
               batch=(B, H, obs, (q_bs, dh), (dk_bs, dh)))

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

    def _conv_trunk_bwd(self, dx, saved, W, inv_scale):
        """dx: fp16 gradient w.r.t. the trunk output [B,7,7,1024]; writes weight/bias grads into the flat buffer."""
        dev = self.dev
        B = dx.shape[0]
        for j in range(4, -1, -1):
            x, y, cols, Hc, cin, cout = saved[j]
            idx = CONV_IDX[j]
            gw, gb = self.grad("%s.meta_net.%d.weight" % (W.mod, idx)), self.grad("%s.meta_net.%d.bias" % (W.mod, idx))
            dy = torch.empty(B, Hc, Hc, cout, device=dev, dtype=F16)
            K.pool_relu_bwd(y, dx, dy, B, Hc, Hc, cout)
            rows = B * Hc * Hc
            if cols is None:
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
