"""CLIP BPE tokenizer (dist_b200/tokenizer.py) against ids produced by the reference's ``dataset/utils/simple_tokenizer.py``
(``tests/golden/tokenizer.json``, written by ``oracle/make_golden_r2.py tokenizer``).  The merge list is data that this repo does not
ship: the golden comparison runs where a copy is reachable (``DISTB200_BPE_PATH``, or the reference checkout of the build container);
the merge algorithm itself is also checked on a hand-made merge list that needs no file."""
import gzip
import json
import os

import pytest
import torch

from conftest import GOLDEN
from dist_b200 import tokenizer as tk

_CANDIDATES = [os.environ.get("DISTB200_BPE_PATH"), "/root/reference/dataset/utils/bpe_simple_vocab_16e6.txt.gz"]
BPE = next((p for p in _CANDIDATES if p and os.path.exists(p)), None)


def test_byte_alphabet_is_a_bijection_with_the_reference_layout():
    a = tk.byte_alphabet()
    assert len(a) == 256 and len(set(a.values())) == 256
    assert a[ord("a")] == "a" and a[ord("!")] == "!" and a[0xAD] != chr(0xAD)
    relocated = [b for b in range(256) if a[b] != chr(b)]
    assert len(relocated) == 68 and [ord(a[b]) for b in relocated] == list(range(256, 256 + 68))      # byte order -> consecutive code points


def test_merge_algorithm_on_a_handmade_list(tmp_path):
    path = tmp_path / "merges.txt.gz"
    merges = ["#version: test", "l o", "lo w</w>", "e r</w>", "n e", "ne w", "l o", "w er</w>"]           # ranks: 'l o' first
    with gzip.open(path, "wb") as f:
        f.write("\n".join(merges).encode())
    t = tk.SimpleTokenizer(str(path), num_merges=len(merges) - 1)
    assert t.merge("low") == ("low</w>",)                           # l o -> lo, then lo w</w>
    assert t.merge("lower") == ("lo", "wer</w>")                    # l o, then e r</w> (rank 2), then w er</w> (rank 6); 'lo w</w>' never applies: w is not word-final
    assert t.merge("newer") == ("new", "er</w>")
    assert t.merge("x") == ("x</w>",)
    ids = t.encode("low newer")
    assert t.decode(ids) == "low newer "
    rows = tk.tokenize(["low", "newer low"], context_length=6, tokenizer=t)
    sot, eot = t.encoder[tk.SOT], t.encoder[tk.EOT]
    assert rows.dtype == torch.int32 and rows.shape == (2, 6)
    assert rows[0].tolist()[:3] == [sot, t.encoder["low</w>"], eot] and rows[0, 3:].sum() == 0
    assert int(rows[1].argmax()) == 4 and rows[1, 4] == eot                     # end-of-text carries the largest id (clip.py:429 pools at argmax)
    with pytest.raises(RuntimeError, match="too long"):
        tk.tokenize(["low low low low low low"], context_length=4, tokenizer=t)
    cut = tk.tokenize(["low low low low low low"], context_length=4, truncate=True, tokenizer=t)
    assert cut[0, -1] == eot and cut[0, 0] == sot


def test_missing_merge_list_is_a_clear_error(monkeypatch):
    monkeypatch.delenv("DISTB200_BPE_PATH", raising=False)
    with pytest.raises(FileNotFoundError, match="DISTB200_BPE_PATH"):
        tk.find_bpe_file(None)


@pytest.mark.skipif(BPE is None, reason="CLIP's bpe_simple_vocab_16e6.txt.gz is not reachable (set DISTB200_BPE_PATH)")
def test_ids_match_the_reference_tokenizer():
    g = json.load(open(os.path.join(GOLDEN, "tokenizer.json")))
    t = tk.SimpleTokenizer(BPE)
    assert len(t.encoder) == g["vocab_size"] == 49408
    got = tk.tokenize(g["prompts"], context_length=77, tokenizer=t)
    assert got.tolist() == g["ids"]
    assert tk.tokenize([g["long_prompt"]], context_length=77, truncate=True, tokenizer=t).tolist() == g["long_truncated"]
    for row, want in zip(g["ids"], g["decoded"]):
        assert t.decode([x for x in row if x != 0]) == want
