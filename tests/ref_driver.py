"""Python mirror of the host-side string/token work of the reference, used to predict what the
unchanged reference binary (oracle/_ref/gliclass_ref_main) feeds the engine:
  * prepare_input  — reference src/preprocessor.c:67-111
  * the stub tokenizer of oracle/stub_tokenizers.c (stand-in for tokenizers-cpp)
  * truncate / pad-to-longest — reference src/tokenizer.c:44-84
"""
import numpy as np


def prepare_input(text, labels, prompt_first):
    lab = "".join("<<LABEL>>" + l.lower() for l in labels)
    return (lab + "<<SEP>>" + text) if prompt_first else (text + lab + "<<SEP>>")


def _fnv1a(b: bytes) -> int:
    h = 2166136261
    for c in b:
        h ^= c
        h = (h * 16777619) & 0xFFFFFFFF
    return h


def stub_encode(s: str, class_token: int, sep_token: int):
    b = s.encode()
    ids, i, n = [1], 0, len(b)
    while i < n:
        if b[i:i + 1] in (b" ", b"\t", b"\n"):
            i += 1
            continue
        if b[i:i + 9] == b"<<LABEL>>":
            ids.append(class_token); i += 9; continue
        if b[i:i + 7] == b"<<SEP>>":
            ids.append(sep_token); i += 7; continue
        j = i
        while j < n and b[j:j + 1] not in (b" ", b"\t", b"\n") and b[j:j + 9] != b"<<LABEL>>" and b[j:j + 7] != b"<<SEP>>":
            j += 1
        ids.append(3 + _fnv1a(b[i:j]) % (class_token - 3))
        i = j
    ids.append(2)
    return ids


def tokenize_batch(strings, class_token, sep_token, max_length=2048):
    enc = [stub_encode(s, class_token, sep_token)[:max_length] for s in strings]
    S = max(len(e) for e in enc)
    ids = np.zeros((len(enc), S), dtype=np.int64)
    mask = np.zeros((len(enc), S), dtype=np.int64)
    for r, e in enumerate(enc):
        ids[r, :len(e)] = e
        mask[r, :len(e)] = 1
    return ids, mask
