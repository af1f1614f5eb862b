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


def test_ascii_fast_path_equals_python_clean():
    """Captions classified printable-ASCII-without-'&' are cleaned inside the library (dc_bpe_tokenize_ex): same ids as the
    Python `clean()` route (simple_tokenizer.py:53-63,126), including stripping, blank runs and upper case."""
    import random
    from declip_b200.tokenizer import SimpleTokenizer, clean
    tok = SimpleTokenizer(MERGES)
    rng = random.Random(0)
    words = "A Photo of the BIG red dog's ball, running  in   park! It's 2023... (two) dogs #tag x_y  ".split(" ")
    caps = ["  " * rng.randint(0, 2) + " ".join(rng.choice(words) for _ in range(rng.randint(1, 40))) + " " * rng.randint(0, 3)
            for _ in range(300)] + ["", " ", "A", "a  b"]
    fast = tok.tokenize(caps, 77, pinned=False)
    for i, c in enumerate(caps):
        assert c.isascii() and c.isprintable() and "&" not in c
        ids = [tok.encoder["<|startoftext|>"]] + tok.encode(c) + [tok.encoder["<|endoftext|>"]]
        if len(ids) > 77:
            ids = ids[:76] + ids[-1:]
        assert fast[i, :len(ids)].tolist() == ids and not fast[i, len(ids):].any(), (i, c)
    # a caption that needs the slow path (entity, non-ASCII, tab) mixed into the same batch
    mixed = ["Tom &amp; Jerry", "café au lait", "tab\tseparated", "plain ascii"]
    out = tok.tokenize(mixed, 77, pinned=False)
    for i, c in enumerate(mixed):
        ids = [tok.encoder["<|startoftext|>"]] + tok.encode(c) + [tok.encoder["<|endoftext|>"]]
        assert out[i, :len(ids)].tolist() == ids


def test_eda_operations():
    """declip_b200.eda restates textaugment.EDA's three operations (declip.py:203-212): seedable, word multiset
    preserved by swap, subset by deletion, synonym table honoured, single words survive."""
    from declip_b200.eda import EDA
    s = "a quick brown fox jumps over the lazy dog near the river bank"
    e1, e2 = EDA(random_state=3), EDA(random_state=3)
    assert [e1.augment(s) for _ in range(20)] == [e2.augment(s) for _ in range(20)]
    e = EDA(random_state=0)
    sw = e.random_swap(s)
    assert sorted(sw.split()) == sorted(s.split()) and sw != s
    dl = e.random_deletion(s, p=0.5)
    assert 0 < len(dl.split()) < len(s.split()) and all(w in s.split() for w in dl.split())
    assert e.random_deletion("word") == "word" and e.random_swap("word") == "word"
    assert e.random_deletion("two words", p=1.0) in ("two", "words")
    assert e.synonym_replacement(s) == s                                   # no synonym source: unchanged
    es = EDA(synonyms={"quick": ["fast"], "the": ["a"]}, random_state=1)
    out = es.synonym_replacement(s, n=5)
    assert "fast" in out.split() and "quick" not in out.split() and out.split().count("the") == 2   # stop words are kept
    assert all(isinstance(x, str) for x in es.augment_batch([s, "x", ""]))
