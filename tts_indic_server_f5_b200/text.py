"""Text front-end of the hot path (host side, pure Python like the reference's).

Mirrors, for Indic scripts, the reference functions the sampler is fed by:
  * `get_tokenizer(path, "custom")`      f5_tts/model/utils.py:124-129
  * `convert_char_to_pinyin`             f5_tts/model/utils.py:140-177
  * `list_str_to_idx`                    f5_tts/model/utils.py:88-95
  * `chunk_text`                         f5_tts/infer/utils_infer.py:61-88
  * duration rule                        f5_tts/infer/utils_infer.py:446-453

The reference tokenises through jieba + pypinyin (not installed here, not needed for Indic text): jieba
emits every non-Han, non-ASCII-alnum character as its own segment, `lazy_pinyin` returns non-Chinese
characters unchanged, so Kannada / Devanagari text becomes ONE TOKEN PER CODE POINT (matras, virama and
ZWNJ included) with spaces preserved.  ASCII alphanumeric runs arrive from jieba as
one segment and get a leading space when the previous token is not one of `" :'\""` (utils.py:156-159);
that rule is reproduced with jieba's documented block regex.  Han text (pinyin conversion) is out of scope
and rejected loudly.
"""
from __future__ import annotations

import re

import torch

_CUSTOM_TRANS = str.maketrans({";": ",", "“": '"', "”": '"', "‘": "'", "’": "'"})  # utils.py:142-144
_ASCII_RUN = re.compile(r"[a-zA-Z0-9]+(?:\.\d+)?%?")
# anything the character scan treats specially: Han, or an ASCII alphanumeric that starts a multi-character _ASCII_RUN match
_NEEDS_SCAN = re.compile(r"[\u3100-\u9fff]|[a-zA-Z0-9](?:[a-zA-Z0-9]|\.\d|%)")


def get_tokenizer(vocab_file: str, tokenizer: str = "custom"):
    """utils.py:124-129: one token per line, line index = id; the trailing newline of every line is
    stripped with `char[:-1]` (so the last line must end with a newline too)."""
    if tokenizer != "custom":
        raise ValueError("only the 'custom' tokenizer (vocab file path) is on the IndicF5 path")
    vocab_char_map = {}
    with open(vocab_file, "r", encoding="utf-8") as f:
        for i, char in enumerate(f):
            vocab_char_map[char[:-1]] = i
    return vocab_char_map, len(vocab_char_map)


def synthetic_indic_vocab() -> list[str]:
    """Synthetic vocab used by benchmarks/tests (the real IndicF5 vocab lives in the HF repo, not in the
    reference tree; the vendored `infer/examples/vocab.txt` has no Indic code points).  Line 0 is " "."""
    toks = [" "]
    toks += [chr(c) for c in range(0x0C80, 0x0D00)]   # Kannada
    toks += [chr(c) for c in range(0x0900, 0x0980)]   # Devanagari
    toks += [".", ",", "!", "?", "\u200c", "\u200d"]  # + ZWNJ / ZWJ
    return toks


def write_vocab(path: str, tokens: list[str]) -> None:
    with open(path, "w", encoding="utf-8") as f:
        for t in tokens:
            f.write(t + "\n")


def _is_han(c: str) -> bool:
    return "\u3100" <= c <= "\u9fff"  # utils.py:146-149


def convert_char_to_pinyin(text_list: list[str], polyphone: bool = True) -> list[list[str]]:
    """utils.py:140-177 restricted to non-Han text: returns one token list per input string."""
    out = []
    for text in text_list:
        text = text.translate(_CUSTOM_TRANS)
        if _NEEDS_SCAN.search(text) is None:      # no Han, no run of >= 2 ASCII alphanumerics (pure Indic text): one token
            out.append(list(text))                # per code point — the scan below would append exactly these
            continue
        chars: list[str] = []
        i = 0
        while i < len(text):
            c = text[i]
            if _is_han(c):
                raise ValueError("Han text needs jieba/pypinyin, which is outside the IndicF5 hot path")
            m = _ASCII_RUN.match(text, i) if c.isascii() and c.isalnum() else None
            if m and len(m.group()) > 1:
                seg = m.group()
                if chars and chars[-1] not in " :'\"":
                    chars.append(" ")
                chars.extend(seg)
                i = m.end()
            else:
                chars.append(c)
                i += 1
        out.append(chars)
    return out


def list_str_to_idx(text: list[str] | list[list[str]], vocab_char_map: dict[str, int], padding_value: int = -1) -> torch.Tensor:
    """utils.py:88-95: unknown -> 0, pad -1."""
    rows = [[vocab_char_map.get(c, 0) for c in t] for t in text]
    nt = max((len(r) for r in rows), default=0)
    ids = torch.full((len(rows), nt), padding_value, dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = torch.tensor(r, dtype=torch.long)
    return ids


def chunk_text(text: str, max_chars: int = 135) -> list[str]:
    """utils_infer.py:61-88: split at sentence punctuation, pack sentences while UTF-8 bytes <= max_chars."""
    chunks, cur = [], ""
    for sentence in re.split(r"(?<=[;:,.!?])\s+|(?<=[；：，。！？])", text):
        piece = sentence + " " if sentence and len(sentence[-1].encode("utf-8")) == 1 else sentence
        if len(cur.encode("utf-8")) + len(sentence.encode("utf-8")) <= max_chars:
            cur += piece
        else:
            if cur:
                chunks.append(cur.strip())
            cur = piece
    if cur:
        chunks.append(cur.strip())
    return chunks


def finish_ref_text(ref_text: str) -> str:
    """utils_infer.py:343-347 and :438-439: sentence-final '. ' rule + single-byte trailing char gets a space."""
    if not ref_text.endswith(". ") and not ref_text.endswith("。"):
        ref_text += " " if ref_text.endswith(".") else ". "
    return ref_text


def estimate_duration(ref_audio_len: int, ref_text: str, gen_text: str, speed: float = 1.0,
                      fix_duration: float | None = None, sample_rate: int = 24000, hop: int = 256) -> int:
    """utils_infer.py:446-453."""
    if fix_duration is not None:
        return int(fix_duration * sample_rate / hop)
    return ref_audio_len + int(ref_audio_len / len(ref_text.encode("utf-8")) * len(gen_text.encode("utf-8")) / speed)


def synthetic_indic_text(n_codepoints: int, seed: int, script: str = "kannada") -> str:
    """Seeded pseudo-text: words of 2-8 code points from the script's block joined by spaces."""
    import random
    rng = random.Random(seed)
    lo, hi = (0x0C85, 0x0CB9) if script == "kannada" else (0x0905, 0x0939)
    signs = list(range(0x0CBE, 0x0CCD)) if script == "kannada" else list(range(0x093E, 0x094D))
    out, total = [], 0
    while total < n_codepoints:
        k = min(rng.randint(2, 8), n_codepoints - total)
        w = "".join(chr(rng.randint(lo, hi)) if rng.random() < 0.65 or j == 0 else chr(rng.choice(signs)) for j in range(k))
        out.append(w)
        total += k + 1
    return " ".join(out)[:n_codepoints]
