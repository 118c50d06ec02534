"""Generate tests/golden/text_tower.npz + text_tokens.json from the REFERENCE's in-tree files (build container only).

  * token ids: perceptor/models/glide_clip/simple_tokenizer.py (imported unmodified; `ftfy`, which is not installed,
    is replaced by an identity `fix_text`) on the reference's own merge table, padded the way open_clip.tokenize /
    SimpleTokenizer.padded_tokens_and_len do (perceptor/models/open_clip.py:99-103).
  * text transformer: perceptor/models/ruclip/model.py CLIP.encode_text (:204-228) on a tiny configuration.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/perceptor")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
PROMPTS = ["a photo of a cat", "Hello, World!  It's 2023 &amp; counting...", "an armchair in the shape of an avocado",
           "ünïcödé señor naïve café", "", "a " * 100, "<|startoftext|>weird<|endoftext|> tokens'll've"]
TINY_TEXT = dict(width=64, heads=2, layers=2, embed=16, context=12, vocab=50)


def main():
    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda t: t
    sys.modules["ftfy"] = ftfy
    sys.path.insert(0, str(REF / "models" / "glide_clip"))
    sys.path.insert(0, str(REF / "models" / "ruclip"))
    import simple_tokenizer as st  # noqa: E402
    import model as ruclip_model  # noqa: E402

    tok = st.SimpleTokenizer()
    rows = []
    for p in PROMPTS:
        ids, _ = tok.padded_tokens_and_len(tok.encode(p), 77)
        rows.append(ids)
    (OUT / "text_tokens.json").write_text(json.dumps({"prompts": PROMPTS, "tokens": rows}))

    from perceptor_b200.text import TextShape, random_text_state_dict

    shape = TextShape(**TINY_TEXT)
    sd = random_text_state_dict(shape, 5)
    g = torch.Generator().manual_seed(6)
    for k in sd:  # non-trivial LayerNorm affine parameters and biases
        if sd[k].dim() == 1:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    eos = shape.vocab - 1
    clip = ruclip_model.CLIP(embed_dim=shape.embed, image_resolution=32, vision_layers=1, vision_width=64,
                             vision_patch_size=8, context_length=shape.context, vocab_size=shape.vocab,
                             transformer_width=shape.width, transformer_heads=shape.heads,
                             transformer_layers=shape.layers, eos_id=eos).eval()
    missing, unexpected = clip.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith("visual.") or m == "logit_scale" for m in missing), (missing, unexpected)
    tokens = torch.randint(1, eos, (5, shape.context), generator=g)
    for i, n in enumerate([3, 12, 7, 1, 9]):
        tokens[i, n - 1] = eos
        tokens[i, n:] = 0
    with torch.no_grad():
        enc = clip.encode_text(tokens)
    np.savez_compressed(OUT / "text_tower.npz", tokens=tokens.numpy(), enc=enc.numpy(),
                        checksum=np.array([float(sum(v.double().sum() for v in sd.values()))]))
    print("text goldens written", enc.shape)




def textoff_fixture():
    """the reference's "text off" target vectors for the ViT towers of the hot path, copied as a
    small fixture so that the GPU box (which has no /root/reference) can use a realistic target."""
    d = json.loads((REF / "losses" / "clip" / "vectors" / "textoff.json").read_text())
    np.savez_compressed(OUT / "textoff_vit.npz", **{k.replace("-", "_"): np.asarray(d[k], dtype=np.float32)
                                                     for k in ("ViT-B-32", "ViT-B-16")})


if __name__ == "__main__":
    main()
    textoff_fixture()
