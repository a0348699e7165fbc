"""Vision expert on the C-ABI kernels: adrefexpert (adrefexpert_v2.py:99-301) = ImageBind-Huge vision trunk
(imagebind_model.py:486-504: ViT-H/14, 32 blocks x 1280, 16 heads of 80, MLP 5120, LayerNorm eps 1e-6, four tapped blocks
7 / 15 / 23 / 31) + the two map heads of `forward`:

  zero-shot (:279-301)  per tapped layer: patch tokens -> image_decoder.fc[i] (1280 -> 1024, :15-28) -> L2-normalise ->
                        100 * tokens . text^T (two prompt-ensemble text embeddings of the sample's class, :68-96) -> softmax at
                        16 x 16 (masks) and after bilinear up-sampling to 224 x 224 (maps); mean over the layers
  k-shot   (:247-278)   per tapped layer: max over the reference images' patches of the cosine similarity; mean over layers;
                        anomaly map = 1 - bilinear(sim), simmask = 1 - sim

The trunk runs on the same kernels as the EVA encoder (patchify + GEMM, LayerNorm, qkv GEMM with the q scale in its epilogue, flash
attention with the head dimension padded 80 -> 96 by TMA zero fill, GEMM + residual, GEMM + erf-GELU); the heads are the row / pixel
kernels of csrc/expert.cu. No arithmetic happens in Python / torch.

Not on the device path (inputs instead): the text tower (24-block CLIP text transformer run once per class name: `text` is the
[n_cls, 2, 1024] table of :68-96, computed offline), image file decoding / CLIP normalisation (data.py), the 90-degree rotation
augmentation of `encode_image_for_one_shot_with_aug` (:172-198, unused by `forward`).

Precision: fp16 GEMM operands, fp32 accumulation, fp32 residual stream and LayerNorm / softmax / cosine statistics (the reference runs
the expert under autocast fp16, adrefexpert_v2.py:246; fp32 streams only move results towards the fp32 oracle).
"""
from dataclasses import dataclass

import torch

from . import kernels as K
from .synthetic import synth

F16, F32 = torch.float16, torch.float32


@dataclass
class ExpertDims:
    img: int = 224
    patch: int = 14
    dim: int = 1280          # imagebind_model.py:493
    depth: int = 32
    heads: int = 16
    mlp_ratio: int = 4       # transformer.py:110,129
    out_layers: tuple = (7, 15, 23, 31)  # imagebind_model.py:74,488-491
    dec_dim: int = 1024      # LinearLayer(1280, 1024, 4), adrefexpert_v2.py:108
    ln_eps: float = 1e-6     # transformer.py:170, imagebind_model.py:311
    out_size: int = 224      # F.interpolate(size=224), adrefexpert_v2.py:276,294

    @property
    def grid(self):
        return self.img // self.patch

    @property
    def tokens(self):
        return self.grid ** 2 + 1

    @property
    def head_dim(self):
        return self.dim // self.heads

    @property
    def mlp_hidden(self):
        return self.mlp_ratio * self.dim


def tiny_expert_dims():
    """Real head width (80), token count (257: the reference preprocessor hard-wires 224 x 224, imagebind_model.py:161-167) and tap
    structure; narrow and shallow otherwise."""
    return ExpertDims(dim=160, depth=4, heads=2, out_layers=(0, 1, 2, 3), dec_dim=64)


VE = "visual_encoder."
_PRE = VE + "modality_preprocessors.vision."
_TRUNK = VE + "modality_trunks.vision."


def expert_state_dict_spec(d: ExpertDims):
    """(key, shape, std, mean) with the names adrefexpert's own state_dict carries (ImageBindModel + image_decoder)."""
    D = d.dim
    S = [(_PRE + "cls_token", (1, 1, D), D ** -0.5, 0), (_PRE + "pos_embedding_helper.pos_embed", (1, d.tokens, D), 0.02, 0),
         (_PRE + "rgbt_stem.proj.1.weight", (D, 3, 2, d.patch, d.patch), 0.02, 0),
         (_TRUNK + "pre_transformer_layer.0.weight", (D,), 0.1, 1.0), (_TRUNK + "pre_transformer_layer.0.bias", (D,), 0.05, 0)]
    for i in range(d.depth):
        b = _TRUNK + "blocks.%d." % i
        S += [(b + "attn.in_proj_weight", (3 * D, D), 0.02, 0), (b + "attn.in_proj_bias", (3 * D,), 0.02, 0),
              (b + "attn.out_proj.weight", (D, D), 0.02 * (2.0 * (i + 1)) ** -0.5, 0), (b + "attn.out_proj.bias", (D,), 0.02, 0),
              (b + "norm_1.weight", (D,), 0.1, 1.0), (b + "norm_1.bias", (D,), 0.05, 0),
              (b + "mlp.fc1.weight", (d.mlp_hidden, D), 0.02, 0), (b + "mlp.fc1.bias", (d.mlp_hidden,), 0.02, 0),
              (b + "mlp.fc2.weight", (D, d.mlp_hidden), 0.02 * (2.0 * (i + 1)) ** -0.5, 0), (b + "mlp.fc2.bias", (D,), 0.02, 0),
              (b + "norm_2.weight", (D,), 0.1, 1.0), (b + "norm_2.bias", (D,), 0.05, 0)]
    for k in range(len(d.out_layers)):
        S += [("image_decoder.fc.%d.weight" % k, (d.dec_dim, D), 0.03, 0), ("image_decoder.fc.%d.bias" % k, (d.dec_dim,), 0.02, 0)]
    return S


def make_expert_state_dict(d: ExpertDims, seed=0, device="cpu"):
    return {k: synth("vision_expert." + k, shape, std, seed, device=device, mean=mean) for k, shape, std, mean in expert_state_dict_spec(d)}


def make_text_features(n_cls, d: ExpertDims, seed=0, device="cpu"):
    """Stand-in for encode_text_with_prompt_ensemble (adrefexpert_v2.py:68-96): [n_cls, 2, dec_dim] unit rows (normal, abnormal)."""
    t = synth("vision_expert.text_features", (n_cls, 2, d.dec_dim), 1.0, seed, device=device, round_fp16=False)
    return t / t.norm(dim=-1, keepdim=True)


class _Obj:
    pass


def _h(t, dev):
    return t.to(device=dev, dtype=F16).contiguous()


def _f(t, dev):
    return t.to(device=dev, dtype=F32).contiguous()


class VisionExpertEngine:
    def __init__(self, sd, dims: ExpertDims, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("VisionExpertEngine needs a CUDA device (there is no CPU path)")
        self.d, self.dev = dims, torch.device(device)
        d, dev = dims, self.dev
        W = _Obj()
        # Conv3d over two identical frames (PadIm2Video "repeat", imagebind_model.py:150-160) == one 14 x 14 conv whose filter is
        # the sum of the two temporal slices
        w3 = sd[_PRE + "rgbt_stem.proj.1.weight"].float()
        kp = 3 * d.patch * d.patch
        W.ldp = (kp + 7) // 8 * 8
        pw = torch.zeros(d.dim, W.ldp, dtype=F16)
        pw[:, :kp] = (w3[:, :, 0] + w3[:, :, 1]).reshape(d.dim, kp).to(F16)
        W.patch_w = pw.to(dev)
        W.cls = _f(sd[_PRE + "cls_token"].reshape(-1), dev)
        W.pos = _f(sd[_PRE + "pos_embedding_helper.pos_embed"].reshape(d.tokens, d.dim), dev)
        W.pre_ln = (_f(sd[_TRUNK + "pre_transformer_layer.0.weight"], dev), _f(sd[_TRUNK + "pre_transformer_layer.0.bias"], dev))
        W.blocks = []
        for i in range(max(d.out_layers) + 1):  # blocks behind the last tap never reach the heads
            b = _TRUNK + "blocks.%d." % i
            B_ = _Obj()
            B_.ln1 = (_f(sd[b + "norm_1.weight"], dev), _f(sd[b + "norm_1.bias"], dev))
            B_.ln2 = (_f(sd[b + "norm_2.weight"], dev), _f(sd[b + "norm_2.bias"], dev))
            B_.wqkv, B_.bqkv = _h(sd[b + "attn.in_proj_weight"], dev), _h(sd[b + "attn.in_proj_bias"], dev)
            B_.wproj, B_.bproj = _h(sd[b + "attn.out_proj.weight"], dev), _h(sd[b + "attn.out_proj.bias"], dev)
            B_.fc1w, B_.fc1b = _h(sd[b + "mlp.fc1.weight"], dev), _h(sd[b + "mlp.fc1.bias"], dev)
            B_.fc2w, B_.fc2b = _h(sd[b + "mlp.fc2.weight"], dev), _h(sd[b + "mlp.fc2.bias"], dev)
            W.blocks.append(B_)
        W.dec = [(_h(sd["image_decoder.fc.%d.weight" % k], dev), _h(sd["image_decoder.fc.%d.bias" % k], dev)) for k in range(len(d.out_layers))]
        self.w = W

    # ------------------------------------------------------------------------------------------------ trunk
    def trunk(self, image, raw=True, unit=False):
        """image fp32 [B, 3, img, img] (device) -> per tapped layer the patch tokens WITHOUT the class token as fp16 [B * P, D]:
        (raw list or None, unit-norm list or None). imagebind_model.py:447-470 + transformer.py:236-283 up to the last tap."""
        d, W, dev = self.d, self.w, self.dev
        B, N, D, H, dh = image.shape[0], d.tokens, d.dim, d.heads, d.head_dim
        T, P = B * N, N - 1
        patches = torch.empty(B * P, W.ldp, device=dev, dtype=F16)
        K.patchify(image, patches, B, 3, d.img, d.patch)
        pe = K.gemm(patches, W.patch_w, out_dtype=F32)  # bias=False, imagebind_model.py:156
        x0 = torch.empty(T, D, device=dev, dtype=F32)
        K.vit_assemble(pe, W.cls, W.pos, x0, B, N, D)  # cat(cls, tokens) + pos_embed, multimodal_preprocessors.py:255-268
        x = torch.empty(T, D, device=dev, dtype=F32)
        h16 = torch.empty(T, D, device=dev, dtype=F16)
        K.norm(x0, W.pre_ln[0], W.pre_ln[1], d.ln_eps, out32=x, out16=h16)  # pre_transformer_layer LayerNorm
        qkv = torch.empty(T, 3 * D, device=dev, dtype=F16)
        ctx = torch.empty(T, D, device=dev, dtype=F16)
        m16 = torch.empty(T, d.mlp_hidden, device=dev, dtype=F16)
        qs, os_ = (3 * D, N * 3 * D, dh), (D, N * D, dh)
        taps, taps_n = ([] if raw else None), ([] if unit else None)
        for i, b in enumerate(W.blocks):
            K.norm(x, b.ln1[0], b.ln1[1], d.ln_eps, out16=h16)
            # nn.MultiheadAttention: (x W^T + b) split into q, k, v; q scaled by dh^-0.5 before q k^T
            K.gemm(h16, b.wqkv, bias=b.bqkv, out=qkv, scale_cols=D, scale=dh ** -0.5)
            K.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], ctx, B, H, N, N, dh, 1.0, qs, qs, qs, os_)
            K.gemm(ctx, b.wproj, bias=b.bproj, res=x, out=x)
            K.norm(x, b.ln2[0], b.ln2[1], d.ln_eps, out16=h16)
            K.gemm(h16, b.fc1w, bias=b.fc1b, act=K.ACT_GELU, out=m16)
            K.gemm(m16, b.fc2w, bias=b.fc2b, res=x, out=x)
            if i in d.out_layers:
                if raw:
                    t = torch.empty(B * P, D, device=dev, dtype=F16)
                    K.expert_tap(x, t, B, N, D, normalize=False)
                    taps.append(t)
                if unit:
                    t = torch.empty(B * P, D, device=dev, dtype=F16)
                    K.expert_tap(x, t, B, N, D, normalize=True)
                    taps_n.append(t)
        return taps, taps_n

    # ------------------------------------------------------------------------------------------------ heads
    def zero_shot_from_taps(self, taps, text):
        """adrefexpert_v2.py:279-301. taps: per layer fp16 [B * P, D]; text fp32 [B, 2, dec_dim] -> (maps [B,1,S,S], masks [B,1,G,G])."""
        d, dev = self.d, self.dev
        P, G = d.tokens - 1, d.grid
        B = taps[0].shape[0] // P
        L = len(taps)
        text = text.to(device=dev, dtype=F32).contiguous()
        assert text.shape == (B, 2, d.dec_dim)
        logits = torch.empty(L, B * P, 2, device=dev, dtype=F32)
        tok = torch.empty(B * P, d.dec_dim, device=dev, dtype=F32)
        for l, t in enumerate(taps):
            K.gemm(t, self.w.dec[l][0], bias=self.w.dec[l][1], out=tok)   # image_decoder.fc[l], :26-27
            K.expert_logits(tok, text, logits[l], B, P, d.dec_dim, 100.0)  # :285-286
        maps = torch.empty(B, 1, d.out_size, d.out_size, device=dev, dtype=F32)
        masks = torch.empty(B, 1, G, G, device=dev, dtype=F32)
        K.expert_maps(logits, maps, masks, L, B, G, d.out_size)  # :287-301
        return maps, masks

    def k_shot_from_taps(self, taps_n, refs_n):
        """adrefexpert_v2.py:264-278. taps_n / refs_n: per layer unit-norm fp16 [B * P, D] / [B * R, D] (R reference patches per
        sample) -> (anomaly maps [B,1,S,S], simmask [B,1,G,G])."""
        d, dev = self.d, self.dev
        P, G, D = d.tokens - 1, d.grid, d.dim
        B = taps_n[0].shape[0] // P
        R = refs_n[0].shape[0] // B
        L = len(taps_n)
        S = torch.empty(B * P, R, device=dev, dtype=F32)
        sim = torch.empty(B * P, device=dev, dtype=F32)
        for l in range(L):
            K.gemm(taps_n[l], refs_n[l], out=S, T=P, F=R, K=D, ldo=R, ldx=D, ldw=D, batch=(B, 1, (P * D, 0), (R * D, 0), (P * R, 0)))
            K.expert_rowmax(S, R, sim, B * P, R, 1.0 / L, accumulate=l > 0)
        maps = torch.empty(B, 1, d.out_size, d.out_size, device=dev, dtype=F32)
        simmask = torch.empty(B, 1, G, G, device=dev, dtype=F32)
        K.expert_sim_maps(sim, maps, simmask, B, G, d.out_size)
        return maps, simmask

    def zero_shot(self, image, text):
        taps, _ = self.trunk(image, raw=True, unit=False)
        return self.zero_shot_from_taps(taps, text)

    def encode_refs(self, ref_images, B):
        """ref_images fp32 [B * k, 3, img, img], the k normal references of sample b at rows [b k, (b + 1) k) (the order
        adrefexpert_v2.py:249-261 stacks them in) -> per layer unit-norm fp16 [B * k * P, D]."""
        assert ref_images.shape[0] % B == 0
        _, refs_n = self.trunk(ref_images, raw=False, unit=True)
        return refs_n

    def k_shot(self, image, ref_images):
        _, tn = self.trunk(image, raw=False, unit=True)
        return self.k_shot_from_taps(tn, self.encode_refs(ref_images, image.shape[0]))

    def both(self, image, text, ref_images=None, refs_n=None):
        """The two calls Myriad makes per batch (myriad.py:342-343) with ONE trunk pass over the query images."""
        taps, tn = self.trunk(image, raw=True, unit=True)
        zs = self.zero_shot_from_taps(taps, text)
        if refs_n is None:
            refs_n = self.encode_refs(ref_images, image.shape[0])
        return zs, self.k_shot_from_taps(tn, refs_n)
