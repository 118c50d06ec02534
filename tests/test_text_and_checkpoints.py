"""§8f-2/3: tokenizer, text tower and checkpoint ingestion (host-side, CPU)."""
import gzip
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from perceptor_b200 import checkpoints, text

GOLDEN = Path(__file__).parent / "golden"
VOCAB = Path(os.environ.get("PCG_BPE_VOCAB", text.DEFAULT_VOCAB))  # the packaged table (perceptor_b200/data/)


def test_byte_alphabet_is_the_gpt2_table():
    table = text.byte_alphabet()
    assert len(table) == 256 and len(set(table.values())) == 256
    assert table[ord("a")] == "a" and table[ord("!")] == "!" and table[0] == chr(256) and table[ord(" ")] == chr(288)
    assert list(table)[:3] == [33, 34, 35] and list(table)[188] == 0


def _write_vocab(tmp_path, merges):
    path = tmp_path / "toy_vocab.txt.gz"
    with gzip.open(path, "wb") as f:
        f.write(("#version: toy\n" + "\n".join(" ".join(m) for m in merges) + "\n").encode())
    return path


def test_bpe_merges_lowest_rank_first_on_a_toy_table(tmp_path):
    merges = [("l", "o"), ("lo", "w</w>"), ("e", "r</w>"), ("n", "e"), ("ne", "w"), ("lo", "w"), ("low", "er</w>")]
    tok = text.SimpleTokenizer(_write_vocab(tmp_path, merges))
    enc = tok.encoder
    assert tok.start_token == 512 + len(merges) and tok.end_token == tok.start_token + 1
    assert tok.encode("low") == [enc["low</w>"]]
    assert tok.encode("Lower  NEW") == [enc["lower</w>"], enc["ne"], enc["w</w>"]]  # cleaned, lower-cased
    # unknown pairs stay split into byte symbols; the end-of-word marker rides on the last symbol
    assert tok.encode("ol") == [enc["o"], enc["l</w>"]]
    assert tok.decode(tok.encode("lower low")) == "lower low "
    rows = text.tokenize(tok, ["low", "low " * 40], context_length=8)
    assert rows.shape == (2, 8) and rows[0].tolist() == [tok.start_token, enc["low</w>"], tok.end_token, 0, 0, 0, 0, 0]
    assert rows[1, 0] == tok.start_token and rows[1, -1] == tok.end_token and (rows[1, 1:-1] == enc["low</w>"]).all()
    with pytest.raises(FileNotFoundError):
        text.SimpleTokenizer(tmp_path / "missing.gz")


def test_token_ids_match_the_reference_tokenizer():
    z = json.loads((GOLDEN / "text_tokens.json").read_text())
    tok = text.SimpleTokenizer(VOCAB)
    assert tok.start_token == 49406 and tok.end_token == 49407
    got = text.tokenize(tok, z["prompts"]).tolist()
    assert got == z["tokens"]


def test_text_tower_matches_reference_encode_text():
    from oracle.make_golden_text import TINY_TEXT

    z = np.load(GOLDEN / "text_tower.npz")
    shape = text.TextShape(**TINY_TEXT)
    sd = text.random_text_state_dict(shape, 5)
    g = torch.Generator().manual_seed(6)
    for k in sd:
        if sd[k].dim() == 1:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(z["checksum"][0])) < 1e-6
    enc = text.encode_text(sd, shape, torch.from_numpy(z["tokens"]), quick_gelu=True, eot_id=shape.vocab - 1)
    assert float((enc - torch.from_numpy(z["enc"])).abs().max()) <= 2e-5


def test_hugging_face_checkpoint_round_trip_against_transformers():
    """An independent implementation as cross-check: a random-init HF CLIPModel's state dict, run
    through normalize_state_dict, must reproduce HF's own image and text embeddings in the oracle ViT / text tower."""
    transformers = pytest.importorskip("transformers")
    from oracle import vit as vit_oracle

    cfg = transformers.CLIPConfig(
        vision_config=dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, image_size=32,
                           patch_size=8, hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=16),
        text_config=dict(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                         max_position_embeddings=12, vocab_size=50, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                         projection_dim=16, eos_token_id=49, bos_token_id=48, pad_token_id=0),
        projection_dim=16)
    torch.manual_seed(0)
    model = transformers.CLIPModel(cfg).eval()
    sd = checkpoints.normalize_state_dict(model.state_dict())
    g = torch.Generator().manual_seed(1)
    pixels = torch.randn(3, 3, 32, 32, generator=g)
    tokens = torch.randint(1, 48, (4, 12), generator=g)
    for i, n in enumerate([3, 12, 6, 2]):
        tokens[i, n - 1] = 49
        tokens[i, n:] = 0
    with torch.no_grad():
        want_img = model.get_image_features(pixel_values=pixels)
        want_txt = model.get_text_features(input_ids=tokens, attention_mask=(tokens != 0).long())
    want_img = getattr(want_img, "pooler_output", want_img)
    want_txt = getattr(want_txt, "pooler_output", want_txt)
    vision = {k: v for k, v in sd.items() if not k.startswith("text.")}
    got_img = vit_oracle.encode(pixels, vision, 8, 2, 2)
    assert float((got_img - want_img).abs().max()) <= 2e-5
    tshape = text.TextShape(width=64, heads=2, layers=2, embed=16, context=12, vocab=50)
    tsd = {k[len("text."):]: v for k, v in sd.items() if k.startswith("text.")}
    assert sorted(tsd) == sorted(text.text_keys(2))
    got_txt = text.encode_text(tsd, tshape, tokens, quick_gelu=True)
    assert float((got_txt - want_txt).abs().max()) <= 2e-5


def test_openai_layout_is_split_without_key_collisions():
    from perceptor_b200.vit import VitShape, random_state_dict, required_keys

    shape = VitShape(image_size=32, patch=8, width=128, layers=2, heads=2, embed=16)
    vis = random_state_dict(shape, 0)
    tshape = text.TextShape(width=64, heads=2, layers=2, embed=16, context=12, vocab=50)
    txt = text.random_text_state_dict(tshape, 1)
    full = {**{"visual." + k: v for k, v in vis.items()}, **txt, "logit_scale": torch.tensor(1.0)}
    out = checkpoints.normalize_state_dict(full)
    assert all(torch.equal(out[k], vis[k]) for k in required_keys(2))
    assert torch.equal(out["text.positional_embedding"], txt["positional_embedding"])  # not the vision table
    assert out["positional_embedding"].shape == vis["positional_embedding"].shape
    assert checkpoints.normalize_state_dict(vis).keys() == vis.keys()
