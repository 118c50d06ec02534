"""Text side of the module API: CLIP byte-pair tokenizer + text transformer, so that
`add_texts_()` works without open_clip.  Runs once per prompt, so it is plain PyTorch on the encoder's device; the
image path is the native one.

Mirrors
  * open_clip.tokenize as called at perceptor/models/open_clip.py:99-103 (77-token context, <|startoftext|> ...
    <|endoftext|>, zero padding, truncation keeps the end token); the algorithm is OpenAI CLIP's byte-level BPE, whose
    in-tree copy is perceptor/models/glide_clip/simple_tokenizer.py:22-164.  `ftfy` is not installed here, so the
    `ftfy.fix_text` call of basic_clean (:60-63) is applied only when ftfy is importable.
  * the text transformer restated at perceptor/models/ruclip/model.py:165-228: token + positional embedding ->
    pre-LN residual attention blocks under a causal mask -> ln_final -> the end-of-text position @ text_projection.
The 48 894 merge rules the tokenizer uses (the head of OpenAI CLIP's published `bpe_simple_vocab_16e6.txt`, MIT
licence) ship as package data (perceptor_b200/data/clip_bpe_merges.txt.gz, see the README there) so that `add_texts_`
works out of the box; `bpe_path=` or PCG_BPE_VOCAB select another table (the full original file works as well).
"""
from __future__ import annotations

import gzip
import html
import os
from dataclasses import dataclass
from functools import lru_cache

import regex as re
import torch
import torch.nn.functional as F

try:  # pragma: no cover - optional dependency of the reference's basic_clean
    import ftfy
except ImportError:  # pragma: no cover
    ftfy = None

DEFAULT_VOCAB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "clip_bpe_merges.txt.gz")
SOT, EOT = "<|startoftext|>", "<|endoftext|>"
N_MERGES = 49152 - 256 - 2
_PATTERN = re.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
                      re.IGNORECASE)


@lru_cache()
def byte_alphabet() -> dict[int, str]:
    """Reversible byte -> printable unicode character map of GPT-2 / CLIP BPE: printable latin-1 bytes map to
    themselves, the remaining 68 bytes to code points 256.. in byte order."""
    keep = [*range(ord("!"), ord("~") + 1), *range(ord("¡"), ord("¬") + 1), *range(ord("®"), ord("ÿ") + 1)]
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # vocabulary order is `keep` first, then the remapped bytes: rebuild the dict in that order
    return {b: table[b] for b in keep + [b for b in range(256) if b not in keep]}


class SimpleTokenizer:
    def __init__(self, bpe_path: str | os.PathLike | None = None):
        bpe_path = bpe_path or os.environ.get("PCG_BPE_VOCAB") or DEFAULT_VOCAB
        if not os.path.exists(bpe_path):
            raise FileNotFoundError(
                f"the CLIP BPE merge table {bpe_path} does not exist: pass bpe_path= or set PCG_BPE_VOCAB (the packaged "
                "copy is perceptor_b200/data/clip_bpe_merges.txt.gz), or add precomputed encodings with "
                "add_encodings_()")
        opener = gzip.open if str(bpe_path).endswith(".gz") else open
        with opener(bpe_path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(line.split()) for line in lines[1:N_MERGES + 1]]
        merges = [m for m in merges if len(m) == 2]
        alphabet = list(byte_alphabet().values())
        vocab = alphabet + [c + "</w>" for c in alphabet] + ["".join(m) for m in merges] + [SOT, EOT]
        self.byte_encoder = byte_alphabet()
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self._cache: dict[str, tuple[str, ...]] = {SOT: (SOT,), EOT: (EOT,)}

    @property
    def start_token(self) -> int:
        return self.encoder[SOT]

    @property
    def end_token(self) -> int:
        return self.encoder[EOT]

    def _bpe(self, token: str) -> tuple[str, ...]:
        """Greedy lowest-rank-first merging of one pre-token (already mapped through the byte alphabet)."""
        hit = self._cache.get(token)
        if hit is not None:
            return hit
        word = [*token[:-1], token[-1] + "</w>"]
        while len(word) > 1:
            best_rank, best = None, None
            for pair in zip(word[:-1], word[1:]):
                r = self.ranks.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best_rank, best = r, pair
            if best is None:
                break
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == best[0] and word[i + 1] == best[1]:
                    merged.append(best[0] + best[1])
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        out = tuple(word)
        self._cache[token] = out
        return out

    @staticmethod
    def clean(text: str) -> str:
        if ftfy is not None:
            text = ftfy.fix_text(text)
        text = html.unescape(html.unescape(text)).strip()
        return re.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text: str) -> list[int]:
        ids: list[int] = []
        for piece in _PATTERN.findall(self.clean(text)):
            mapped = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self._bpe(mapped))
        return ids

    def decode(self, ids) -> str:
        inverse = {c: b for b, c in self.byte_encoder.items()}
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(inverse[c] for c in text).decode("utf-8", errors="replace").replace("</w>", " ")


def tokenize(tokenizer: SimpleTokenizer, texts, context_length: int = 77) -> torch.Tensor:
    """int64 [len(texts), context_length]: <sot> tokens <eot> then zeros; over-long prompts are cut and end in <eot>."""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, text in enumerate(texts):
        ids = [tokenizer.start_token, *tokenizer.encode(text), tokenizer.end_token]
        if len(ids) > context_length:
            ids = ids[:context_length]
            ids[-1] = tokenizer.end_token
        out[i, : len(ids)] = torch.tensor(ids)
    return out


@dataclass(frozen=True)
class TextShape:
    width: int
    heads: int
    layers: int
    embed: int
    context: int = 77
    vocab: int = 49408


# OpenAI / open_clip text towers that pair with the vision towers of vit.SHAPES
TEXT_SHAPES = {
    "ViT-B-32": TextShape(512, 8, 12, 512), "ViT-B-16": TextShape(512, 8, 12, 512),
    "ViT-L-14": TextShape(768, 12, 12, 768), "ViT-L-14-336": TextShape(768, 12, 12, 768),
    "ViT-H-14": TextShape(1024, 16, 24, 1024), "ViT-g-14": TextShape(1024, 16, 24, 1024),  # open_clip model_configs
}


def text_keys(layers: int) -> list[str]:
    keys = ["token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection"]
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        keys += [p + s for s in ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias",
                                 "attn.out_proj.weight", "attn.out_proj.bias", "ln_2.weight", "ln_2.bias",
                                 "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias")]
    return keys


def random_text_state_dict(shape: TextShape, seed: int = 0) -> dict[str, torch.Tensor]:
    """Random init with the scales of perceptor/models/ruclip/model.py:230-263 (no checkpoints offline)."""
    g = torch.Generator().manual_seed(seed)
    w, n = shape.width, shape.layers
    proj_std, attn_std, fc_std = (w**-0.5) * ((2 * n) ** -0.5), w**-0.5, (2 * w) ** -0.5

    def rnd(*size, std):
        return torch.randn(*size, generator=g) * std

    sd = {"token_embedding.weight": rnd(shape.vocab, w, std=0.02), "positional_embedding": rnd(shape.context, w, std=0.01),
          "ln_final.weight": torch.ones(w), "ln_final.bias": torch.zeros(w), "text_projection": rnd(w, shape.embed, std=w**-0.5)}
    for i in range(n):
        p = f"transformer.resblocks.{i}."
        sd.update({p + "ln_1.weight": torch.ones(w), p + "ln_1.bias": torch.zeros(w),
                   p + "ln_2.weight": torch.ones(w), p + "ln_2.bias": torch.zeros(w),
                   p + "attn.in_proj_weight": rnd(3 * w, w, std=attn_std), p + "attn.in_proj_bias": torch.zeros(3 * w),
                   p + "attn.out_proj.weight": rnd(w, w, std=proj_std), p + "attn.out_proj.bias": torch.zeros(w),
                   p + "mlp.c_fc.weight": rnd(4 * w, w, std=fc_std), p + "mlp.c_fc.bias": torch.zeros(4 * w),
                   p + "mlp.c_proj.weight": rnd(w, 4 * w, std=proj_std), p + "mlp.c_proj.bias": torch.zeros(w)})
    return sd


def encode_text(sd: dict[str, torch.Tensor], shape: TextShape, tokens: torch.Tensor, quick_gelu: bool = True,
                eot_id: int | None = None) -> torch.Tensor:
    """[n, context] token ids -> [n, embed] (un-normalised), float32 on the device of the weights."""
    dev = sd["positional_embedding"].device
    tokens = tokens.to(dev)
    n, ctx = tokens.shape
    x = F.embedding(tokens, sd["token_embedding.weight"]) + sd["positional_embedding"][:ctx]
    hd = shape.width // shape.heads
    for i in range(shape.layers):
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (shape.width,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        q, k, v = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]).view(n, ctx, 3, shape.heads, hd) \
            .permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(n, ctx, shape.width)
        x = x + F.linear(a, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = F.layer_norm(x, (shape.width,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        h = h * torch.sigmoid(1.702 * h) if quick_gelu else F.gelu(h)
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    x = F.layer_norm(x, (shape.width,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    # the end-of-text token has the largest id of the vocabulary (open_clip: argmax; ruclip: == eos_id)
    pos = tokens.argmax(dim=-1) if eot_id is None else (tokens == eot_id).float().argmax(dim=-1)
    return x[torch.arange(n, device=dev), pos] @ sd["text_projection"]
