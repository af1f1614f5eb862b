"""C++ BPE tokenizer (csrc/bpe.cu via declip_b200.tokenizer) against ids produced by the UNMODIFIED reference
SimpleTokenizer / TextTransformer.tokenize (tools/make_golden_bpe.py -> tests/golden/bpe_tokenizer.json): bit-exact ids for
contractions, digits, punctuation runs, HTML entities, accents, CJK, emoji, special-token literals, empty and over-long
captions; truncation and padding at two context lengths.  Host code only — no GPU needed."""
import gzip
import json
import os

import pytest
import torch

from declip_b200.tokenizer import SimpleTokenizer, clean

HERE = os.path.dirname(__file__)
G = json.load(open(os.path.join(HERE, "golden", "bpe_tokenizer.json")))
MERGES = os.path.join(HERE, "golden", "bpe_small_merges.txt")


@pytest.fixture(scope="module")
def tok():
    return SimpleTokenizer(MERGES, threads=4)


def test_vocabulary_layout(tok):
    assert len(tok.encoder) == G["vocab"] == G["merges"] + 515 + 1      # + the '' entry of the empty last line
    for k, v in G["special"].items():
        assert tok.encoder[k] == v
    assert "<|mask|>" in tok.encoder and "definitely-not-a-token" not in tok.encoder


def test_encode_matches_reference(tok):
    for text, want in zip(G["texts"], G["encode"]):
        assert tok.encode(text) == want, (text, clean(text))


@pytest.mark.parametrize("L", [77, 12])
def test_tokenize_matches_reference(tok, L):
    ids, lengths = tok.tokenize(G["texts"], context_length=L, return_length=True)
    ref = G["tokenize_%d" % L]
    assert ids.dtype == torch.long and tuple(ids.shape) == (len(G["texts"]), L)
    assert ids.tolist() == ref["ids"]
    assert lengths.tolist() == ref["lengths"]


def test_threads_do_not_change_the_result():
    a = SimpleTokenizer(MERGES, threads=1).tokenize(G["texts"] * 7, 77)
    b = SimpleTokenizer(MERGES, threads=8).tokenize(G["texts"] * 7, 77)
    assert torch.equal(a, b)
    assert SimpleTokenizer(MERGES).tokenize([], 77).shape == (0, 77)


def test_full_size_table_ids(tmp_path):
    """The 48 894-merge table of the model goldens (dummy merges that never apply): vocabulary 49 409, <|mask|> = 49406,
    SOT = 49407, EOT = 49408 (simple_tokenizer.py:66-75) and byte-level ids identical to the reference."""
    path = tmp_path / "fake_bpe.txt.gz"
    with gzip.open(path, "wt") as f:
        f.write("#version: fake\n")
        for i in range(49152 - 256 - 2):
            f.write("a%d b%d\n" % (i, i))
    t = SimpleTokenizer(str(path))
    assert len(t.encoder) == G["fake_vocab"] == 49409
    for k, v in G["fake_special"].items():
        assert t.encoder[k] == v
    for text, want in zip(G["texts"][:6], G["fake_encode"]):
        assert t.encode(text) == want


def test_text_transformer_uses_the_tokenizer():
    from declip_b200.model.text_transformer import text_transformers
    m = text_transformers(embed_dim=512, transformer_layers=1, bpe_path=MERGES, text_encode_type="Transformer")
    ids = m.tokenize(G["texts"][:4], context_length=77)
    assert ids.tolist() == G["tokenize_77"]["ids"][:4]
