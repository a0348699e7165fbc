"""Host-side orchestration of the Myriad hot path over the C-ABI kernels (no arithmetic happens in Python/torch).

MyriadEngine owns device copies of the weights (prepared once: fp16 GEMM operands, fused/concatenated
projections, NHWC conv filters) and sequences kernel launches for

  vit_forward        eva_vit.py:324-340          (a1-a6 of SURVEY.md §8a)
  encode_img         myriad.py:241-272 / 274-306 (a7-a16)
  llama prefill      modeling_llama.py:466-716   (a18-a26, forward)
  generate           myriad.py:433-454 + greedy search + conversation.py:96-107 (a27, a29)

Precision contract (stated tolerance of the parity tests): fp16 tensor-core operands, fp32 accumulation, fp32
residual streams / softmax / norm statistics. The reference's CUDA path keeps fp16 residuals in the ViT and
Q-Former (autocast) — keeping them in fp32 only moves results closer to the fp32 CPU oracle.
"""
import math
import os

import torch

from . import kernels as K
from .synthetic import CONV_CHANNELS, CONV_IDX, MyriadDims

F16, F32 = torch.float16, torch.float32


def _h(t, device):
    return t.to(device=device, dtype=F16).contiguous()


def _f(t, device):
    return t.to(device=device, dtype=F32).contiguous()


class _Obj:
    pass


def l_hidden_ok(dims):
    """The small-batch kernel of csrc/gemv.cu (and its RMSNorm hand-over) needs K % 128 == 0 for every LLaMA projection."""
    return dims.llama.hidden % 128 == 0 and dims.llama.inter % 128 == 0


class MyriadEngine:
    # Inference layout of the LLaMA weights: LoRA A rows appended to the fused qkv weight (the rank-8 update is applied
    # inside the RoPE kernel) and gate/up rows interleaved in blocks of 64 for the fused SwiGLU GEMM epilogue. The trainer
    # keeps the plain layout (it needs the pre-activation gate/up values and xa for the backward pass).
    FUSED_LLAMA = True

    def __init__(self, sd, dims: MyriadDims, device="cuda:0", max_batch=8, max_seq=512):
        self.d = dims
        self.dev = torch.device(device)
        self.max_batch, self.max_seq = max_batch, max_seq
        self._prep_vit(sd)
        self._prep_experts(sd)
        self._prep_qformer(sd)
        self._prep_llama(sd)
        self._decode_graphs = {}
        # measurement hook (bench.py): when set to a list, every greedy_decode appends (event before the first decode step, event
        # after the last one, number of decode steps launched), so the decode-step time is taken inside the timed region
        self.decode_timing = None
        # MYR_MEGA=1: one persistent kernel per decode step (decode_mega.cu) instead of ~230 graph-captured launches. Measured
        # on B200: 4.11 ms/step against 3.70 ms/step for the PDL-chained multi-kernel step, so it is opt-in;
        # tests/test_engine_gpu.py keeps the two paths bit-identical.
        import os
        self.use_mega = os.environ.get("MYR_MEGA", "0") == "1"
        # MYR_GEMV=0 sends T <= 4 GEMMs back to the tcgen05 kernel (which has no fused-norm prologue)
        self.fuse_small_batch_norm = self.FUSED_LLAMA and os.environ.get("MYR_GEMV", "1") != "0" and l_hidden_ok(dims)

    # ------------------------------------------------------------------------------------------ weights
    def _prep_vit(self, sd):
        v, dev = self.d.vit, self.dev
        p = "visual_encoder."
        W = _Obj()
        kp = 3 * v.patch * v.patch
        W.ldp = (kp + 7) // 8 * 8
        pw = torch.zeros(v.dim, W.ldp, dtype=F16)
        pw[:, :kp] = sd[p + "patch_embed.proj.weight"].reshape(v.dim, kp).to(F16)
        W.patch_w, W.patch_b = pw.to(dev), _h(sd[p + "patch_embed.proj.bias"], dev)
        W.cls, W.pos = _f(sd[p + "cls_token"].reshape(-1), dev), _f(sd[p + "pos_embed"].reshape(v.tokens, v.dim), dev)
        W.blocks = []
        for i in range(v.depth):
            b = p + "blocks.%d." % i
            B_ = _Obj()
            B_.ln1 = (_f(sd[b + "norm1.weight"], dev), _f(sd[b + "norm1.bias"], dev))
            B_.ln2 = (_f(sd[b + "norm2.weight"], dev), _f(sd[b + "norm2.bias"], dev))
            B_.wqkv = _h(sd[b + "attn.qkv.weight"], dev)
            qb, vb = sd[b + "attn.q_bias"], sd[b + "attn.v_bias"]
            B_.bqkv = _h(torch.cat([qb, torch.zeros_like(vb), vb]), dev)  # eva_vit.py:120-124: K has no bias
            B_.wproj, B_.bproj = _h(sd[b + "attn.proj.weight"], dev), _h(sd[b + "attn.proj.bias"], dev)
            B_.fc1w, B_.fc1b = _h(sd[b + "mlp.fc1.weight"], dev), _h(sd[b + "mlp.fc1.bias"], dev)
            B_.fc2w, B_.fc2b = _h(sd[b + "mlp.fc2.weight"], dev), _h(sd[b + "mlp.fc2.bias"], dev)
            W.blocks.append(B_)
        W.ln_vision = (_f(sd["ln_vision.weight"], dev), _f(sd["ln_vision.bias"], dev))
        W.ad1, W.ad2 = _f(sd["expert_adaptor.conv1.weight"], dev), _f(sd["expert_adaptor.conv2.weight"], dev)
        self.vitw = W

    def _prep_conv(self, sd, mod, head_k):
        dev = self.dev
        W = _Obj()
        W.direct, W.gemm = [], []
        for j, idx in enumerate(CONV_IDX):
            w, b = sd["%s.meta_net.%d.weight" % (mod, idx)], sd["%s.meta_net.%d.bias" % (mod, idx)]
            wn = w.permute(0, 2, 3, 1).contiguous()  # [Cout, kh, kw, Cin]
            if j < 3:
                W.direct.append((_f(wn, dev), _f(b, dev), CONV_CHANNELS[j], CONV_CHANNELS[j + 1]))
            else:
                W.gemm.append((_h(wn.reshape(w.shape[0], -1), dev), _h(b, dev), CONV_CHANNELS[j], CONV_CHANNELS[j + 1]))
        w, b = sd["%s.meta_net.15.weight" % mod], sd["%s.meta_net.15.bias" % mod]
        W.head_w, W.head_b, W.head_k = _h(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), dev), _h(b, dev), head_k
        return W

    def _prep_experts(self, sd):
        self.instw = self._prep_conv(sd, "VEInstructor", 1) if self.d.use_instructor else None
        self.tokw = self._prep_conv(sd, "VETokenizer", 5) if self.d.use_tokenizer else None
        if self.tokw is not None:
            self.tokw.base_prompts = _f(sd["VETokenizer.base_prompts"], self.dev)

    def refresh_trainables(self, sd):
        """Re-read the TRAINABLE tensors (LoraAdaptorV2, both conv stacks + base_prompts, LoRA A / B) from `sd` into the prepared
        device copies, in place: captured decode graphs and cached tensor maps keep pointing at live storage. Called by the
        drop-in model after optimizer steps (the frozen ViT / Q-Former / LLaMA operands are never touched)."""
        with torch.no_grad():
            W = self.vitw
            W.ad1.copy_(sd["expert_adaptor.conv1.weight"])
            W.ad2.copy_(sd["expert_adaptor.conv2.weight"])
            for mod, cur, head_k in (("VEInstructor", self.instw, 1), ("VETokenizer", self.tokw, 5)):
                if cur is None:
                    continue
                new = self._prep_conv(sd, mod, head_k)
                for (a, ab, _, _), (b, bb, _, _) in zip(cur.direct + cur.gemm, new.direct + new.gemm):
                    a.copy_(b)
                    ab.copy_(bb)
                cur.head_w.copy_(new.head_w)
                cur.head_b.copy_(new.head_b)
            if self.tokw is not None:
                self.tokw.base_prompts.copy_(sd["VETokenizer.base_prompts"])
            r, D = self.d.lora_r, self.d.llama.hidden
            if r > 0:
                for i, L in enumerate(self.llw.layers):
                    pl = "llama_model.base_model.model.model.layers.%d.self_attn." % i
                    a = torch.cat([sd[pl + "q_proj.lora_A.default.weight"], sd[pl + "v_proj.lora_A.default.weight"]])
                    bq, bv = sd[pl + "q_proj.lora_B.default.weight"], sd[pl + "v_proj.lora_B.default.weight"]
                    if self.FUSED_LLAMA:
                        L.wqkv[3 * D:3 * D + 2 * r].copy_(a)
                        L.lora.bq.copy_(bq)
                        L.lora.bv.copy_(bv)
                    else:
                        L.lora.a.copy_(a)
                        L.lora.bq.copy_(bq * L.lora.scale)
                        L.lora.bv.copy_(bv * L.lora.scale)

    def _prep_qformer(self, sd):
        q, dev = self.d.qf, self.dev
        p = "Qformer.bert."
        W = _Obj()
        W.query_tokens = _f(sd["query_tokens"].reshape(q.num_query, q.hidden), dev)
        W.emb_ln = (_f(sd[p + "embeddings.LayerNorm.weight"], dev), _f(sd[p + "embeddings.LayerNorm.bias"], dev))
        W.layers, ckv_w, ckv_b = [], [], []
        for i in range(q.layers):
            lp = p + "encoder.layer.%d." % i
            L = _Obj()
            a = lp + "attention."
            L.wqkv = _h(torch.cat([sd[a + "self.query.weight"], sd[a + "self.key.weight"], sd[a + "self.value.weight"]]), dev)
            L.bqkv = _h(torch.cat([sd[a + "self.query.bias"], sd[a + "self.key.bias"], sd[a + "self.value.bias"]]), dev)
            L.wo, L.bo = _h(sd[a + "output.dense.weight"], dev), _h(sd[a + "output.dense.bias"], dev)
            L.ln_a = (_f(sd[a + "output.LayerNorm.weight"], dev), _f(sd[a + "output.LayerNorm.bias"], dev))
            L.cross = i % q.cross_freq == 0
            if L.cross:
                c = lp + "crossattention."
                L.cq_w, L.cq_b = _h(sd[c + "self.query.weight"], dev), _h(sd[c + "self.query.bias"], dev)
                L.ckv_index = len(ckv_w)
                ckv_w.append(torch.cat([sd[c + "self.key.weight"], sd[c + "self.value.weight"]]))
                ckv_b.append(torch.cat([sd[c + "self.key.bias"], sd[c + "self.value.bias"]]))
                L.co_w, L.co_b = _h(sd[c + "output.dense.weight"], dev), _h(sd[c + "output.dense.bias"], dev)
                L.ln_c = (_f(sd[c + "output.LayerNorm.weight"], dev), _f(sd[c + "output.LayerNorm.bias"], dev))
            L.fi_w, L.fi_b = _h(sd[lp + "intermediate_query.dense.weight"], dev), _h(sd[lp + "intermediate_query.dense.bias"], dev)
            L.fo_w, L.fo_b = _h(sd[lp + "output_query.dense.weight"], dev), _h(sd[lp + "output_query.dense.bias"], dev)
            L.ln_f = (_f(sd[lp + "output_query.LayerNorm.weight"], dev), _f(sd[lp + "output_query.LayerNorm.bias"], dev))
            W.layers.append(L)
        # every cross-attention layer's K/V projection reads the same image features: one GEMM for all of them
        W.ckv_w, W.ckv_b = _h(torch.cat(ckv_w), dev), _h(torch.cat(ckv_b), dev)
        W.proj_w, W.proj_b = _h(sd["llama_proj.weight"], dev), _h(sd["llama_proj.bias"], dev)
        self.qfw = W

    def _prep_llama(self, sd):
        l, dev = self.d.llama, self.dev
        p = "llama_model.model."
        W = _Obj()
        W.embed = _h(sd[p + "embed_tokens.weight"], dev)
        W.layers = []
        for i in range(l.layers):
            lp = p + "layers.%d." % i
            L = _Obj()
            qkv_rows = [sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.v_proj.weight"]]
            pl = "llama_model.base_model.model.model.layers.%d.self_attn." % i
            if self.FUSED_LLAMA and self.d.lora_r > 0:
                qkv_rows += [sd[pl + "q_proj.lora_A.default.weight"], sd[pl + "v_proj.lora_A.default.weight"]]
            L.wqkv = _h(torch.cat([t.to(dev) for t in qkv_rows]), dev)  # frozen rows may sit on the host, trainable LoRA rows on the device
            del qkv_rows
            L.wo = _h(sd[lp + "self_attn.o_proj.weight"], dev)
            g_, u_ = _h(sd[lp + "mlp.gate_proj.weight"], dev), _h(sd[lp + "mlp.up_proj.weight"], dev)
            if self.FUSED_LLAMA:
                assert l.inter % 64 == 0
                L.wgu = torch.stack([g_.reshape(l.inter // 64, 64, l.hidden), u_.reshape(l.inter // 64, 64, l.hidden)],
                                    1).reshape(2 * l.inter, l.hidden).contiguous()
            else:
                L.wgu = torch.cat([g_, u_])
            del g_, u_
            L.wd = _h(sd[lp + "mlp.down_proj.weight"], dev)
            L.n1, L.n2 = _f(sd[lp + "input_layernorm.weight"], dev), _f(sd[lp + "post_attention_layernorm.weight"], dev)
            L.lora = None
            if self.d.lora_r > 0:
                s = self.d.lora_alpha / self.d.lora_r
                L.lora = _Obj()
                L.lora.scale = s
                if self.FUSED_LLAMA:
                    assert self.d.lora_r == 8, "the fused RoPE + LoRA kernel is written for the reference's r = 8 (myriad.py:172)"
                    L.lora.bq = _h(sd[pl + "q_proj.lora_B.default.weight"], dev)
                    L.lora.bv = _h(sd[pl + "v_proj.lora_B.default.weight"], dev)
                else:
                    L.lora.a = _h(torch.cat([sd[pl + "q_proj.lora_A.default.weight"].to(dev), sd[pl + "v_proj.lora_A.default.weight"].to(dev)]), dev)
                    L.lora.bq = _h(sd[pl + "q_proj.lora_B.default.weight"] * s, dev)
                    L.lora.bv = _h(sd[pl + "v_proj.lora_B.default.weight"] * s, dev)
            W.layers.append(L)
        W.norm = _f(sd[p + "norm.weight"], dev)
        W.lm_head = _h(sd["llama_model.lm_head.weight"], dev)
        half = l.head_dim // 2
        inv = 1.0 / (10000.0 ** (torch.arange(0, l.head_dim, 2).float() / l.head_dim))  # modeling_llama.py:80
        fr = torch.outer(torch.arange(l.max_pos).float(), inv)
        W.cos, W.sin = fr.cos().to(dev).contiguous(), fr.sin().to(dev).contiguous()
        assert W.cos.shape == (l.max_pos, half)
        self.llw = W
        self.kcache = self.vcache = None

    # ---------------------------------------------------------------------------------------------- ViT
    def vit_forward(self, image):
        """image fp32 [B,3,img,img] on device -> fp32 [B*N, D] residual stream after the last block."""
        v, W, dev = self.d.vit, self.vitw, self.dev
        B, N, D, H, dh = image.shape[0], v.tokens, v.dim, v.heads, v.head_dim
        T = B * N
        patches = torch.empty(B * (N - 1), W.ldp, device=dev, dtype=F16)
        K.patchify(image, patches, B, 3, v.img, v.patch)
        pe = K.gemm(patches, W.patch_w, bias=W.patch_b, out_dtype=F32)
        x = torch.empty(T, D, device=dev, dtype=F32)
        K.vit_assemble(pe, W.cls, W.pos, x, B, N, D)
        h16 = torch.empty(T, D, device=dev, dtype=F16)
        qkv = torch.empty(T, 3 * D, device=dev, dtype=F16)
        ctx = torch.empty(T, D, device=dev, dtype=F16)
        m16 = torch.empty(T, v.mlp_hidden, device=dev, dtype=F16)
        qs, os_ = (3 * D, N * 3 * D, dh), (D, N * D, dh)
        for b in W.blocks:
            K.norm(x, b.ln1[0], b.ln1[1], v.ln_eps, out16=h16)
            K.gemm(h16, b.wqkv, bias=b.bqkv, out=qkv, scale_cols=D, scale=dh ** -0.5)  # q * scale, eva_vit.py:128
            K.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], ctx, B, H, N, N, dh, 1.0, qs, qs, qs, os_)
            K.gemm(ctx, b.wproj, bias=b.bproj, res=x, out=x)
            K.norm(x, b.ln2[0], b.ln2[1], v.ln_eps, out16=h16)
            K.gemm(h16, b.fc1w, bias=b.fc1b, act=K.ACT_GELU, out=m16)
            K.gemm(m16, b.fc2w, bias=b.fc2b, res=x, out=x)
        return x

    # ------------------------------------------------------------------------------------ expert tokens
    def _conv_trunk(self, maps, W):
        """maps fp32 [B,1,224,224] -> NHWC fp16 [B,7,7,1024] (networks.py:98-122 / 159-182)."""
        dev = self.dev
        B, Hc = maps.shape[0], maps.shape[2]
        x = maps
        for wn, b, cin, cout in W.direct:
            y = torch.empty(B, Hc // 2, Hc // 2, cout, device=dev, dtype=F16)
            K.conv3x3_relu_pool(x, wn, b, y, B, Hc, Hc, cin, cout)
            x, Hc = y, Hc // 2
        for wg, b, cin, cout in W.gemm:
            cols = torch.empty(B * Hc * Hc, 9 * cin, device=dev, dtype=F16)
            K.im2col(x, cols, B, Hc, Hc, cin, 3, 3, 1)
            y = K.gemm(cols, wg, bias=b, act=K.ACT_RELU)
            x = torch.empty(B, Hc // 2, Hc // 2, cout, device=dev, dtype=F16)
            K.maxpool2(y, x, B, Hc, Hc, cout)
            Hc //= 2
        return x

    def ve_instructor(self, maps, out, group_stride):
        """VEInstructorV2.forward networks.py:149-153 -> 49 fp32 rows per sample written at out (ld = 768)."""
        B = maps.shape[0]
        trunk = self._conv_trunk(maps, self.instw)
        K.gemm(trunk.reshape(B * 49, 1024), self.instw.head_w, bias=self.instw.head_b, out=out, out_group_rows=49,
               out_group_stride=group_stride, T=B * 49, ldo=self.d.qf.hidden)

    def ve_tokenizer(self, maps, out, group_stride):
        """VETokenizer.forward networks.py:191-197 -> 9 learned prompts + 9 conv tokens per sample (fp32)."""
        B, W = maps.shape[0], self.tokw
        trunk = self._conv_trunk(maps, W)
        cols = torch.empty(B * 9, 25 * 1024, device=self.dev, dtype=F16)
        K.im2col(trunk, cols, B, 7, 7, 1024, 5, 5, 0)
        Dl = self.d.llama.hidden
        K.copy_rows(W.base_prompts, out, B, 9, Dl, Dl, 0, Dl, group_stride)
        K.gemm(cols, W.head_w, bias=W.head_b, out=out[9 * Dl:], out_group_rows=9, out_group_stride=group_stride, T=B * 9,
               ldo=Dl)

    # ----------------------------------------------------------------------------------------- Q-Former
    def _qf_attn_out(self, ctx, w, b, ln, h32, tmp32, h16, eps):
        K.gemm(ctx, w, bias=b, res=h32, out=tmp32)         # BertSelfOutput: dense + residual ...
        K.norm(tmp32, ln[0], ln[1], eps, out16=h16, out32=h32)  # ... + LayerNorm (Qformer.py:278-289)

    def qformer_forward(self, q32, enc16, B, Q):
        """q32 fp32 [B*Q, hidden] query embeds (consumed), enc16 fp16 [B*N, D_vit] -> (h32, h16) last hidden state."""
        q, W, dev = self.d.qf, self.qfw, self.dev
        Hd, H = q.hidden, q.heads
        dh = Hd // H
        N = self.d.vit.tokens
        T = B * Q
        h32, h16 = q32, torch.empty(T, Hd, device=dev, dtype=F16)
        tmp32 = torch.empty(T, Hd, device=dev, dtype=F32)
        qkv = torch.empty(T, 3 * Hd, device=dev, dtype=F16)
        ctx = torch.empty(T, Hd, device=dev, dtype=F16)
        cq = torch.empty(T, Hd, device=dev, dtype=F16)
        ff = torch.empty(T, q.inter, device=dev, dtype=F16)
        K.norm(q32, W.emb_ln[0], W.emb_ln[1], q.ln_eps, out16=h16, out32=h32)
        ckv = K.gemm(enc16, W.ckv_w, bias=W.ckv_b)  # [B*N, n_cross * 2 * hidden]
        ldkv = ckv.shape[1]
        scale = 1.0 / math.sqrt(dh)
        for L in W.layers:
            K.gemm(h16, L.wqkv, bias=L.bqkv, out=qkv)
            s = (3 * Hd, Q * 3 * Hd, dh)
            K.attention(qkv, qkv[:, Hd:], qkv[:, 2 * Hd:], ctx, B, H, Q, Q, dh, scale, s, s, s, (Hd, Q * Hd, dh))
            self._qf_attn_out(ctx, L.wo, L.bo, L.ln_a, h32, tmp32, h16, q.ln_eps)
            if L.cross:
                K.gemm(h16, L.cq_w, bias=L.cq_b, out=cq)
                kk = ckv[:, L.ckv_index * 2 * Hd:]
                ks = (ldkv, N * ldkv, dh)
                K.attention(cq, kk, kk[:, Hd:], ctx, B, H, Q, N, dh, scale, (Hd, Q * Hd, dh), ks, ks, (Hd, Q * Hd, dh))
                self._qf_attn_out(ctx, L.co_w, L.co_b, L.ln_c, h32, tmp32, h16, q.ln_eps)
            K.gemm(h16, L.fi_w, bias=L.fi_b, act=K.ACT_GELU, out=ff)
            K.gemm(ff, L.fo_w, bias=L.fo_b, res=h32, out=tmp32)
            K.norm(tmp32, L.ln_f[0], L.ln_f[1], q.ln_eps, out16=h16, out32=h32)
        return h32, h16

    # --------------------------------------------------------------------------------------- encode_img
    def num_image_tokens(self, stage):
        nq = self.d.qf.num_query + (49 if stage in (1, 2) else 0)
        return nq + (18 if stage in (0, 1) else 0)

    def encode_img(self, image, maps, stage, out=None, out_batch_stride=None):
        """Myriad.encode_img myriad.py:241-272 (== encode_img_oneshot :274-306 given the one-shot maps).
        Writes fp32 [B, n_tokens, llama_hidden] (optionally into a slice of a larger [B, L, hidden] buffer)."""
        d, dev = self.d, self.dev
        B, N, Dv = image.shape[0], d.vit.tokens, d.vit.dim
        Hq, Dl = d.qf.hidden, d.llama.hidden
        x = self.vit_forward(image)
        enc16 = torch.empty(B * N, Dv, device=dev, dtype=F16)
        W = self.vitw
        K.norm(x, W.ln_vision[0], W.ln_vision[1], 1e-5, out16=enc16, w1=W.ad1, w2=W.ad2)  # ln_vision(expert_adaptor(x))
        nq0 = d.qf.num_query
        Q = nq0 + (49 if stage in (1, 2) else 0)
        q32 = torch.empty(B * Q, Hq, device=dev, dtype=F32)
        K.copy_rows(self.qfw.query_tokens, q32, B, nq0, Hq, Hq, 0, Hq, Q * Hq)
        if stage in (1, 2):
            self.ve_instructor(maps, q32[nq0:], Q * Hq)
        _, h16 = self.qformer_forward(q32, enc16, B, Q)
        n_tok = self.num_image_tokens(stage)
        if out is None:
            out = torch.empty(B, n_tok, Dl, device=dev, dtype=F32)
            out_batch_stride = n_tok * Dl
        flat = out.reshape(-1)
        K.gemm(h16, self.qfw.proj_w, bias=self.qfw.proj_b, out=flat, out_group_rows=Q, out_group_stride=out_batch_stride,
               T=B * Q, ldo=Dl)
        if stage in (0, 1):
            self.ve_tokenizer(maps, flat[Q * Dl:], out_batch_stride)
        return out

    # -------------------------------------------------------------------------------------------- LLaMA
    def _ensure_cache(self, B, S):
        l = self.d.llama
        if self.kcache is None or self.kcache.shape[1] < B or self.kcache.shape[2] < S:
            Bm, Sm = max(B, self.max_batch), max(S, self.max_seq)
            self.kcache = torch.zeros(l.layers, Bm, Sm, l.hidden, device=self.dev, dtype=F16)
            self.vcache = torch.zeros_like(self.kcache)
            self._decode_graphs = {}

    def _llama_layer(self, L, li, h32, bufs, B, S, pos, kv_len, cache_off, cache_off_dev, Skv, causal, hand=None, next_gamma=None,
                     kv_cap=0, split_ws=None):
        """LlamaDecoderLayer.forward modeling_llama.py:247-299 as 8 launches: RMSNorm -> qkv (+ LoRA A rows) GEMM -> RoPE +
        LoRA B + KV-cache append -> flash attention -> o_proj GEMM (+ residual) -> RMSNorm -> gate/up GEMM with fused
        SwiGLU -> down GEMM (+ residual). Weights are static, so each GEMM may prefetch them under the previous kernel."""
        l = self.d.llama
        D, H, dh, T = l.hidden, l.heads, l.head_dim, B * S
        x16, qkv, ctx, _, act = bufs
        kc, vc = self.kcache[li], self.vcache[li]
        ldq = qkv.shape[1]
        # T <= 4 (greedy decode at the reference's batch sizes): the small-batch weight-streaming kernels hand the RMSNorm over
        # between themselves, so a layer is 5 dependent launches instead of 7
        fuse = T <= 4 and self.fuse_small_batch_norm
        # decode steps also hand the RMSNorm over between launches: o_proj / down_proj write rn_f16(h * gamma_next) and their
        # slice's sum of squares next to the fp32 stream, the next projection copies those rows and only applies the scale
        hand = hand if fuse else None
        if hand is not None and li > 0:
            K.gemm(hand.yb, L.wqkv, out=qkv, w_static=True, norm_ss=(hand.ssb, l.eps))
        else:
            K.norm(h32, L.n1, None, l.eps, rms=True, out16=x16)
            K.gemm(x16, L.wqkv, out=qkv, w_static=True)
        lora = (L.lora.bq, L.lora.bv, self.d.lora_r, L.lora.scale) if L.lora is not None else None  # myriad.py:171-178
        if S == 1 and dh == 128 and kv_len is not None:
            # decode: rotary + LoRA-B + cache append + attention over the cache in one CUDA-core launch
            nxt = self.kcache.stride(0) if li + 1 < l.layers else 0
            K.decode_attention(qkv, B, H, dh, pos, self.llw.cos, self.llw.sin, kc, vc, kv_len, ctx, 1.0 / math.sqrt(dh),
                               cache_off=cache_off, cache_off_dev=cache_off_dev, lora=lora, next_layer_stride=nxt, kv_cap=kv_cap,
                               split_ws=split_ws)
        else:
            K.rope_cache(qkv, B, S, H, dh, pos, self.llw.cos, self.llw.sin, kc, vc, cache_off=cache_off,
                         cache_off_dev=cache_off_dev, lora=lora)
            cs = (kc.stride(1), kc.stride(0), dh)
            K.attention(qkv, kc, vc, ctx, B, H, S, Skv, dh, 1.0 / math.sqrt(dh), (ldq, S * ldq, dh), cs, cs, (D, S * D, dh),
                        causal=causal, q_off=0, kv_len=kv_len)
        if hand is not None:
            K.gemm(ctx, L.wo, res=h32, out=h32, w_static=True, post_norm=(L.n2, hand.ya, hand.ssa))
            K.gemm(hand.ya, L.wgu, act=K.ACT_SWIGLU, out=act, w_static=True, norm_ss=(hand.ssa, l.eps))
            K.gemm(act, L.wd, res=h32, out=h32, w_static=True, post_norm=(next_gamma, hand.yb, hand.ssb))
            return
        K.gemm(ctx, L.wo, res=h32, out=h32, w_static=True)
        K.norm(h32, L.n2, None, l.eps, rms=True, out16=x16)
        K.gemm(x16, L.wgu, act=K.ACT_SWIGLU, out=act, w_static=True)
        K.gemm(act, L.wd, res=h32, out=h32, w_static=True)

    def _llama_bufs(self, T):
        l, dev = self.d.llama, self.dev
        wq = 3 * l.hidden + (2 * self.d.lora_r if (self.FUSED_LLAMA and self.d.lora_r) else 0)
        gu = None if self.FUSED_LLAMA else torch.empty(T, 2 * l.inter, device=dev, dtype=F16)
        return (torch.empty(T, l.hidden, device=dev, dtype=F16), torch.empty(T, wq, device=dev, dtype=F16),
                torch.empty(T, l.hidden, device=dev, dtype=F16), gu, torch.empty(T, l.inter, device=dev, dtype=F16))

    def llama_prefill(self, embeds32, kv_len=None, all_logits=False):
        """embeds32 fp32 [B, S, D] (consumed as the residual stream). kv_len int32 [B] = valid (unpadded) length per
        row (right padding, myriad.py:395-404) or None. Fills the KV cache slots [0, S). Returns fp32 logits
        [B, V] of the last position (generate) or [B, S, V] (training forward, modeling_llama.py:690)."""
        l, dev = self.d.llama, self.dev
        B, S, D = embeds32.shape
        self._ensure_cache(B, S)
        h32 = embeds32.reshape(B * S, D)
        pos = torch.arange(S, device=dev, dtype=torch.int32).repeat(B)  # modeling_llama.py:510-517
        bufs = self._llama_bufs(B * S)
        for li, L in enumerate(self.llw.layers):
            self._llama_layer(L, li, h32, bufs, B, S, pos, kv_len, 0, None, S, True)
        if all_logits:
            x16 = bufs[0]
            K.norm(h32, self.llw.norm, None, l.eps, rms=True, out16=x16)
            return K.gemm(x16, self.llw.lm_head, out_dtype=F32, w_static=True).reshape(B, S, l.vocab)
        last = h32.reshape(B, S, D)[:, S - 1]
        x16 = torch.empty(B, D, device=dev, dtype=F16)
        K.norm(last, self.llw.norm, None, l.eps, rms=True, out16=x16)
        return K.gemm(x16, self.llw.lm_head, out_dtype=F32, w_static=True)

    # ------------------------------------------------------------------------------------------- decode
    def _decode_step(self, st):
        """One token for every row: embed cur_tok -> 32 layers against the cache -> logits -> greedy bookkeeping.
        Every step-dependent scalar is read from the device state, so the launch sequence is graph-replayable."""
        l = self.d.llama
        B = st.B
        K.embed(self.llw.embed, st.cur_tok, st.h32)
        hand = st.hand if (B <= 4 and self.fuse_small_batch_norm) else None
        for li, L in enumerate(self.llw.layers):
            nxt = self.llw.layers[li + 1].n1 if li + 1 < l.layers else self.llw.norm
            self._llama_layer(L, li, st.h32, st.bufs, B, 1, st.pos, st.kv_len, 0, st.cache_off, st.Skv, False, hand=hand, next_gamma=nxt,
                              kv_cap=st.kv_cap if hand is not None else 0, split_ws=st.attn_ws)
        if hand is not None:
            K.gemm(hand.yb, self.llw.lm_head, out=st.logits, w_static=True, norm_ss=(hand.ssb, l.eps))
        else:
            K.norm(st.h32, self.llw.norm, None, l.eps, rms=True, out16=st.bufs[0])
            K.gemm(st.bufs[0], self.llw.lm_head, out=st.logits, w_static=True)
        K.greedy_step(st.logits, st.state, st.scratch, B, l.vocab, st.max_new, st.min_new, l.eos, st.stops, st.n_stops,
                      st.stop_len)

    def _mega_plan(self, st):
        """Op list of one decode step for the persistent kernel (kernels.MegaPlan): embedding gather + first RMSNorm, then
        per layer qkv(+LoRA A) GEMM -> attention -> o_proj (+residual, tail RMSNorm) -> gate/up + SwiGLU -> down_proj
        (+residual, tail RMSNorm of the next layer / final norm), and lm_head. Same buffers and weights as _decode_step."""
        l, W = self.d.llama, self.llw
        B = st.B
        x16, qkv, ctx, _, act = st.bufs
        ops = [K.mega_embed(W.embed, st.cur_tok, st.h32, norm=(st.h32, x16, W.layers[0].n1, l.eps))]
        for li, L in enumerate(W.layers):
            kc, vc = self.kcache[li], self.vcache[li]
            lora = (L.lora.bq, L.lora.bv, self.d.lora_r, L.lora.scale) if L.lora is not None else None
            nxt = W.layers[li + 1].n1 if li + 1 < len(W.layers) else W.norm
            ops += [K.mega_gemm(x16, L.wqkv, qkv, K.MEGA_F16),
                    K.mega_attn(qkv, B, l.heads, l.head_dim, st.pos, W.cos, W.sin, kc, vc, st.kv_len, ctx, 1.0 / math.sqrt(l.head_dim),
                                st.cache_off, lora=lora),
                    K.mega_gemm(ctx, L.wo, st.h32, K.MEGA_RES32, norm=(st.h32, x16, L.n2, l.eps)),
                    K.mega_gemm(x16, L.wgu, act, K.MEGA_SWIGLU),
                    K.mega_gemm(act, L.wd, st.h32, K.MEGA_RES32, norm=(st.h32, x16, nxt, l.eps))]
        ops.append(K.mega_gemm(x16, W.lm_head, st.logits, K.MEGA_F32))
        return K.MegaPlan(ops, self.dev)

    def _mega_ok(self, B, Skv):
        l = self.d.llama
        return (self.FUSED_LLAMA and self.use_mega and B <= 8 and l.head_dim == 128 and Skv <= 4096 and l.hidden % 8 == 0
                and (self.d.lora_r in (0, 8)))

    def _decode_step_mega(self, st):
        l = self.d.llama
        st.mega.launch()
        K.greedy_step(st.logits, st.state, st.scratch, st.B, l.vocab, st.max_new, st.min_new, l.eos, st.stops, st.n_stops,
                      st.stop_len)

    def _decode_state(self, B, max_new, min_new, stop_seqs, Skv, kv_cap=0):
        l, dev = self.d.llama, self.dev
        st = _Obj()
        st.B, st.max_new, st.min_new, st.Skv = B, max_new, min_new, Skv
        st.kv_cap = kv_cap  # upper bound on the visible cache over this state's decode steps (0: unbounded)
        st.state = torch.zeros(4 + 4 * B + B * max_new, device=dev, dtype=torch.int32)
        st.unfinished = st.state[4:4 + B]
        st.cur_tok = st.state[4 + B:4 + 2 * B]
        st.kv_len = st.state[4 + 2 * B:4 + 3 * B]
        st.pos = st.state[4 + 3 * B:4 + 4 * B]
        st.cache_off = st.state[2:3]
        st.scratch = torch.zeros(B, device=dev, dtype=torch.int32)
        st.stop_len = max(len(s) for s in stop_seqs) if stop_seqs else 1
        st.n_stops = len(stop_seqs)
        stops = torch.full((max(st.n_stops, 1), st.stop_len), -1, dtype=torch.int32)
        for i, s in enumerate(stop_seqs):
            stops[i, :len(s)] = torch.tensor(s, dtype=torch.int32)
        st.stops = stops.to(dev)
        st.h32 = torch.empty(B, l.hidden, device=dev, dtype=F32)
        st.bufs = self._llama_bufs(B)
        st.logits = torch.empty(B, l.vocab, device=dev, dtype=F32)
        st.hand = _Obj()  # RMSNorm hand-over buffers of the small-batch path (o_proj -> gate/up, down_proj -> next qkv / lm_head)
        st.hand.ya, st.hand.yb = (torch.zeros(B, l.hidden, device=dev, dtype=F16) for _ in range(2))
        st.hand.ssa, st.hand.ssb = (torch.zeros(K.NORM_SS_FLOATS, device=dev, dtype=F32) for _ in range(2))
        # long caches: decode attention runs one CTA per 128-key chunk and combines the partials here (zeroed once; the kernel
        # leaves its arrival counters at zero)
        st.attn_ws = None
        if Skv > 256 and os.environ.get("MYR_ATTN_SPLIT", "1") != "0":
            st.attn_ws = torch.zeros(K.decode_attn_split_bytes(B, l.heads, Skv), device=dev, dtype=torch.uint8)
        st.graph = None
        st.mega = self._mega_plan(st) if self._mega_ok(B, Skv) else None
        return st

    def greedy_decode(self, embeds32, max_new_tokens=90, stop_seqs=((835,), (2277, 29937)), min_new_tokens=1,
                      use_graph=True, sync_every=8):
        """Greedy search from inputs_embeds (all-ones attention mask, as Myriad.generate passes none):
        prefill, then one token per step. Returns int64 [B, n_new] NEW tokens only (CPU tensor).
        The stop decision is taken on the device; the host reads the (step, done) flags only every `sync_every` graph
        replays, so the GPU runs decode steps back to back. Replays issued after the stop leave the state untouched."""
        dev = self.dev
        B, S, _ = embeds32.shape
        self._ensure_cache(B, S + max_new_tokens)
        Skv = self.kcache.shape[2]
        # decode attention keeps K / V of caches up to 256 tokens in shared memory; the bound is baked into the captured graph
        kv_cap = S + max_new_tokens if (S + max_new_tokens <= 256 and os.environ.get("MYR_ATTN_TMA", "1") != "0") else 0
        key = (B, max_new_tokens, min_new_tokens, tuple(map(tuple, stop_seqs)), Skv, kv_cap)
        st = self._decode_graphs.get(key)
        if st is None:
            st = self._decode_state(B, max_new_tokens, min_new_tokens, stop_seqs, Skv, kv_cap)
            self._decode_graphs[key] = st
        init = torch.zeros(4 + 4 * B, dtype=torch.int32)
        init[2] = S - 1
        init[4:4 + B] = 1
        init[4 + 2 * B:4 + 3 * B] = S
        init[4 + 3 * B:4 + 4 * B] = S - 1
        st.state[:4 + 4 * B].copy_(init.to(dev, non_blocking=True))
        logits = self.llama_prefill(embeds32)
        l = self.d.llama
        K.greedy_step(logits, st.state, st.scratch, B, l.vocab, st.max_new, st.min_new, l.eos, st.stops, st.n_stops, st.stop_len)
        flags = st.state[:2]
        ev0 = n_replays = None
        if self.decode_timing is not None and st.graph is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            n_replays = 0
        while True:
            step, done = flags.tolist()
            if done:
                break
            if use_graph:
                step_fn = self._decode_step_mega if st.mega is not None else self._decode_step
                if st.graph is None:
                    snap = st.state.clone()
                    step_fn(st)  # warm-up (sets function attributes, sizes the split-K workspace)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    st.state.copy_(snap)
                    n0 = K.launch_count()
                    with torch.cuda.graph(g):
                        step_fn(st)
                    st.graph_nodes = K.launch_count() - n0
                    st.graph = g
                    st.state.copy_(snap)  # capture does not execute; replay from the snapshot
                for _ in range(max(1, min(sync_every, max_new_tokens - step))):
                    st.graph.replay()
                    K.note_graph_replay(st.graph_nodes)
                    if n_replays is not None:
                        n_replays += 1
            else:
                self._decode_step(st)
        if ev0 is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self.decode_timing.append((ev0, ev1, n_replays))
        n = int(st.state[0].item())
        toks = st.state[4 + 4 * B:].reshape(B, st.max_new)[:, :n]
        return toks.to(torch.int64).cpu()

    # ----------------------------------------------------------------------------------------- generate
    def build_inputs_embeds(self, image, maps, stage, ids_before, ids_after, with_bos=False, text_ids=None):
        """prompt_wrap myriad.py:354-375 (+ bos / target-text embeds of Myriad.forward :395-421): writes
        [bos] + before + image tokens + after [+ text] straight into one fp32 [B, L, D] buffer.
        ids_before / ids_after: int64 [n] (shared by the batch) or [B, n] (per sample; equal lengths, as the
        reference's torch.stack at myriad.py:371 requires)."""
        l, dev = self.d.llama, self.dev
        B, D = image.shape[0], l.hidden
        n_img = self.num_image_tokens(stage)
        ids_before, ids_after = ids_before.long().cpu(), ids_after.long().cpu()
        shared = ids_before.dim() == 1
        if shared:
            ids_before, ids_after = ids_before[None], ids_after[None]
        G = ids_before.shape[0]
        assert G in (1, B) and ids_after.shape[0] == G
        nb, na = ids_before.shape[1], ids_after.shape[1]
        n0 = 1 if with_bos else 0
        nt = 0 if text_ids is None else text_ids.shape[1]
        L = n0 + nb + n_img + na + nt
        buf = torch.empty(B, L, D, device=dev, dtype=F32)
        flat = buf.reshape(-1)
        head = torch.cat([torch.full((G, n0), l.bos, dtype=torch.long), ids_before, ids_after], 1)  # [G, n0+nb+na]
        nh = head.shape[1]
        tmp = torch.empty(G * nh, D, device=dev, dtype=F32)
        K.embed(self.llw.embed, head.reshape(-1).to(dev), tmp)
        gs = 0 if G == 1 else nh * D
        if n0 + nb:
            K.copy_rows(tmp, flat, B, n0 + nb, D, D, gs, D, L * D)
        if na:
            K.copy_rows(tmp[n0 + nb:], flat[(n0 + nb + n_img) * D:], B, na, D, D, gs, D, L * D)
        self.encode_img(image, maps, stage, out=flat[(n0 + nb) * D:], out_batch_stride=L * D)
        if nt:
            tt = torch.empty(B * nt, D, device=dev, dtype=F32)
            K.embed(self.llw.embed, text_ids.reshape(-1).to(dev).long(), tt)
            K.copy_rows(tt, flat[(n0 + nb + n_img + na) * D:], B, nt, D, D, nt * D, D, L * D)
        return buf

    def generate(self, image, maps, ids_before, ids_after, max_new_tokens=90, stop_seqs=((835,), (2277, 29937)),
                 min_new_tokens=1, use_graph=True):
        """Myriad.generate myriad.py:433-454: stage fixed to 1, no bos, greedy search. -> int64 [B, n_new] (CPU)."""
        emb = self.build_inputs_embeds(image, maps, 1, ids_before, ids_after)
        return self.greedy_decode(emb, max_new_tokens, stop_seqs, min_new_tokens, use_graph)
