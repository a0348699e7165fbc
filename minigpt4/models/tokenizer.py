"""Tokenizer plumbing. The reference uses `LlamaTokenizer.from_pretrained(llama_model, use_fast=False)` with
pad = eos (myriad.py:181-182). When that directory exists we do the same; offline (no Vicuna files) a deterministic
stand-in with the same call surface is used so the hot path can be exercised end to end on synthetic weights."""
import os
import zlib

import torch


class _Batch(dict):
    def __getattr__(self, k):
        return self[k]

    def to(self, device):
        return _Batch({k: v.to(device) for k, v in self.items()})


class SyntheticLlamaTokenizer:
    """Whitespace/punctuation pieces hashed into [3, vocab). Stable across processes; NOT a language tokenizer."""

    bos_token_id, eos_token_id, unk_token_id = 1, 2, 0
    bos_token, eos_token = "<s>", "</s>"

    def __init__(self, vocab_size=32000):
        self.vocab_size = vocab_size
        self.pad_token = self.eos_token
        self.padding_side = "right"
        self._seen = {}  # id -> piece for everything encoded so far, so decode() returns readable text for known pieces

    @property
    def pad_token_id(self):
        return self.eos_token_id

    def _pieces(self, text):
        out, cur = [], ""
        for ch in text:
            if ch.isalnum():
                cur += ch
            else:
                if cur:
                    out.append(cur)
                    cur = ""
                if not ch.isspace():
                    out.append(ch)
        if cur:
            out.append(cur)
        return out

    def encode(self, text, add_special_tokens=True):
        # '###' is the stop marker of the conversation template (conversation.py:128-130): keep its real Vicuna id
        ids = []
        for p in self._pieces(text.replace("###", " \x00 ")):
            if p == "\x00" and self.vocab_size > 835:
                ids.append(835)
                continue
            i = 3 + zlib.crc32(p.encode()) % (self.vocab_size - 3)
            self._seen.setdefault(i, p)
            ids.append(i)
        return ([self.bos_token_id] if add_special_tokens else []) + ids

    def __call__(self, text, return_tensors=None, add_special_tokens=True, padding=False, truncation=False, max_length=None,
                 **_):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [self.encode(t, add_special_tokens) for t in texts]
        if truncation and max_length is not None:
            rows = [r[:max_length] for r in rows]
        L = max(len(r) for r in rows) if rows else 0
        ids = torch.full((len(rows), L), self.pad_token_id, dtype=torch.long)
        mask = torch.zeros((len(rows), L), dtype=torch.long)
        for i, r in enumerate(rows):
            if self.padding_side == "right":
                ids[i, :len(r)] = torch.tensor(r, dtype=torch.long)
                mask[i, :len(r)] = 1
            else:
                ids[i, L - len(r):] = torch.tensor(r, dtype=torch.long)
                mask[i, L - len(r):] = 1
        return _Batch(input_ids=ids, attention_mask=mask)

    def decode(self, ids, skip_special_tokens=False, **_):
        ids = ids.tolist() if torch.is_tensor(ids) else list(ids)
        toks = []
        for i in ids:
            if skip_special_tokens and i in (0, 1, 2):
                continue
            toks.append("###" if i == 835 else self._seen.get(i, "<%d>" % i))
        return " ".join(toks)

    def batch_decode(self, batch, **kw):
        return [self.decode(r, **kw) for r in batch]


def load_llama_tokenizer(path, vocab_size=32000, allow_synthetic=False):
    """The Vicuna SentencePiece tokenizer of `path` (myriad.py:181-182). The hash-based stand-in is returned ONLY when the
    caller runs on synthetic weights (`allow_synthetic`: MYRIAD_SYNTHETIC_WEIGHTS=1 or an explicit `weights=` mapping): with a
    real checkpoint a missing / mistyped tokenizer path must fail, not silently produce garbage prompts and decodes."""
    if path and os.path.isdir(path):
        if os.path.exists(os.path.join(path, "tokenizer.model")):
            from transformers import LlamaTokenizer
            tok = LlamaTokenizer.from_pretrained(path, use_fast=False)
            tok.pad_token = tok.eos_token
            return tok
        if os.path.exists(os.path.join(path, "tokenizer.json")):
            from transformers import AutoTokenizer
            tok = AutoTokenizer.from_pretrained(path)
            tok.pad_token = tok.eos_token
            return tok
    if allow_synthetic:
        return SyntheticLlamaTokenizer(vocab_size)
    raise FileNotFoundError("no LLaMA tokenizer (tokenizer.model / tokenizer.json) under llama_model=%r; the synthetic tokenizer is "
                            "only used with MYRIAD_SYNTHETIC_WEIGHTS=1 or weights=" % (path,))
