"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/charades_queries.npz from the data files the reference ships
("sample real lines from Charades_sta_train.txt (held-out: _test.txt), tokenise by whitespace after
stripping '.', map through Charades_word2id.json, lengths 2-10; embedding initialised from data/glove_weights"):

    glove   [1302, 300] f32   data/glove_weights (row 0 = padding; main.py:94 copies it into query_encoder.embedding.weight)
    train_tokens / test_tokens  [N, 10] int16, zero padded;  train_len / test_len  [N] int8
The first 4096 train / 1024 test queries whose words are all in the vocabulary (dataset.py:90-94 uses nltk.word_tokenize, which
is not installed here; lines it would split differently -- possessives, hyphens -- are skipped).  Needs /root/reference; CPU only.

    python oracle/make_query_fixture.py"""
import json
import os

import numpy as np
import torch

REF = os.environ.get("DRN_REFERENCE_ROOT", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tokens(split, limit, word2id):
    toks, lens = [], []
    for line in open(os.path.join(REF, "data/dataset/Charades/Charades_sta_%s.txt" % split)):
        _, sent = line.strip().split("##")
        words = sent.replace(".", "").split()
        if not (2 <= len(words) <= 10) or any(w not in word2id for w in words):
            continue
        row = np.zeros(10, dtype=np.int16)
        row[:len(words)] = [word2id[w] for w in words]
        toks.append(row)
        lens.append(len(words))
        if len(toks) == limit:
            break
    return np.stack(toks), np.array(lens, dtype=np.int8)


def main():
    word2id = json.load(open(os.path.join(REF, "data/dataset/Charades/Charades_word2id.json")))
    glove = torch.load(os.path.join(REF, "data/glove_weights")).float().numpy()
    assert glove.shape == (1302, 300)
    tr_t, tr_l = tokens("train", 4096, word2id)
    te_t, te_l = tokens("test", 1024, word2id)
    out = os.path.join(REPO, "tests", "golden", "charades_queries.npz")
    np.savez_compressed(out, glove=glove, train_tokens=tr_t, train_len=tr_l, test_tokens=te_t, test_len=te_l)
    print("wrote", out, os.path.getsize(out), "bytes;", tr_t.shape, te_t.shape, "mean length %.2f" % tr_l.mean())


if __name__ == "__main__":
    main()
