"""CLIP's byte-level BPE tokenizer (reference: ``dataset/utils/simple_tokenizer.py:64-179``, itself OpenAI CLIP's).

Turns label prompts into the int token ids ``[C, 77]`` that ``CLIP.encode_text`` takes (``clip.py:419-434``): text is cleaned,
lower-cased and split by CLIP's pattern; every piece is spelled in a reversible byte alphabet and merged bottom-up by the ranked
merge list; ``<|startoftext|>`` / ``<|endoftext|>`` frame the sequence, the rest of the context is zero.  The end-of-text token
has the largest id, which is what the text tower's ``argmax`` pooling relies on (``clip.py:429``).

The merge list (``bpe_simple_vocab_16e6.txt.gz``, 1.3 MB, published with CLIP) is data, not code, and is NOT shipped here: pass its
path, or set ``DISTB200_BPE_PATH``.  ``ftfy`` (mojibake repair in the reference's ``basic_clean``) is used when installed;
without it the text passes through unchanged, which is identical for well-formed input such as dataset label lists.
"""

import gzip
import html
import os

import torch

try:  # the pattern needs Unicode property classes; `regex` is what the reference uses
    import regex as re
except ImportError:  # pragma: no cover
    re = None

try:
    import ftfy
except ImportError:
    ftfy = None

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
_NUM_MERGES = 49152 - 256 - 2            # lines of the merge file that CLIP's 49 408-entry vocabulary uses (simple_tokenizer.py:68)
_PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""


def byte_alphabet():
    """Byte value -> printable stand-in character.  The 188 bytes that already are visible latin-1 characters stand for
    themselves; the remaining 68 (controls, space, soft hyphen ...) are moved to code points 256, 257, ... in byte order."""
    visible = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, spare = {}, 256
    for b in range(256):
        if b in visible:
            table[b] = chr(b)
        else:
            table[b] = chr(spare)
            spare += 1
    return table


def find_bpe_file(path=None):
    for cand in (path, os.environ.get("DISTB200_BPE_PATH")):
        if cand:
            if not os.path.exists(cand):
                raise FileNotFoundError(cand)
            return cand
    raise FileNotFoundError("CLIP's merge list bpe_simple_vocab_16e6.txt.gz is not shipped with dist_b200: pass bpe_path= or set "
                            "DISTB200_BPE_PATH (the reference keeps it at dataset/utils/bpe_simple_vocab_16e6.txt.gz)")


class SimpleTokenizer:
    def __init__(self, bpe_path=None, num_merges=_NUM_MERGES):
        if re is None:
            raise ImportError("the CLIP tokenizer needs the `regex` package (Unicode property classes)")
        opener = gzip.open if str(find_bpe_file(bpe_path)).endswith(".gz") else open
        with opener(find_bpe_file(bpe_path), "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(line.split()) for line in lines[1:1 + num_merges]]          # line 0 is a version header
        merges = [m for m in merges if len(m) == 2]
        self.alphabet = byte_alphabet()
        self.byte_of = {c: b for b, c in self.alphabet.items()}
        # vocabulary order fixes the ids: byte symbols in the order of the reference's table (visible bytes first, then the
        # relocated ones), the same symbols as word-final variants, one entry per merge, the two specials
        ordered = [b for b in range(256) if self.alphabet[b] == chr(b)] + [b for b in range(256) if self.alphabet[b] != chr(b)]
        symbols = [self.alphabet[b] for b in ordered]
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.decoder = {i: tok for tok, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.pattern = re.compile(_PATTERN, re.IGNORECASE)
        self._memo = {SOT: (SOT,), EOT: (EOT,)}

    # ---- text -> pieces ------------------------------------------------------------------------
    @staticmethod
    def clean(text):
        if ftfy is not None:
            text = ftfy.fix_text(text)
        text = html.unescape(html.unescape(text)).strip()
        return re.sub(r"\s+", " ", text).strip().lower()

    def merge(self, piece):
        """Symbols of one pre-token after applying the ranked merges: always the adjacent pair with the best rank, every
        occurrence left to right, until no listed pair is left."""
        hit = self._memo.get(piece)
        if hit is not None:
            return hit
        word = list(piece[:-1]) + [piece[-1] + "</w>"]
        while len(word) > 1:
            best = min(zip(word, word[1:]), key=lambda p: self.rank.get(p, float("inf")))
            if best not in self.rank:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = out
        res = tuple(word)
        self._memo[piece] = res
        return res

    def encode(self, text):
        ids = []
        for piece in self.pattern.findall(self.clean(text)):
            spelled = "".join(self.alphabet[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self.merge(spelled))
        return ids

    def decode(self, tokens):
        text = "".join(self.decoder[int(t)] for t in tokens)
        raw = bytearray(self.byte_of[c] for c in text)              # the word-final marker is spelled with stand-alone characters
        return raw.decode("utf-8", errors="replace").replace("</w>", " ")


_DEFAULT = None


def tokenize(texts, context_length=77, truncate=False, tokenizer=None):
    """``tokenize`` of the reference (``simple_tokenizer.py:135-179``): ``[n, context_length]`` int32 ids, zero padded; a prompt
    that does not fit raises unless ``truncate`` (then the last kept token becomes end-of-text)."""
    global _DEFAULT
    if tokenizer is None:
        if _DEFAULT is None:
            _DEFAULT = SimpleTokenizer()
        tokenizer = _DEFAULT
    if isinstance(texts, str):
        texts = [texts]
    sot, eot = tokenizer.encoder[SOT], tokenizer.encoder[EOT]
    out = torch.zeros(len(texts), context_length, dtype=torch.int32)
    for i, text in enumerate(texts):
        ids = [sot] + tokenizer.encode(text) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError("Input {} is too long for context length {}".format(text, context_length))
            ids = ids[:context_length]
            ids[-1] = eot
        out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out
