"""Turns the reference's packaged data (`/root/reference/data/*.rda`) into the
fixture `packaged_data.npz` committed next to this script.  Run here (where /root/reference exists):

    python tests/golden/make_packaged_data.py

Contents (all in the reference's state order a,t,c,g — `dimnames` of the packaged matrices):
  within   float64 [10][3][4][4]   withinClusTransProbs[[i]][[r]][row, col]   (row-stochastic, P[parent, child])
  between  float64 [10][3][4][4]   betweenClusTransProbs
  Q        float64 [4][4]          QmatrixPosada2001
  seqs     uint8   [6][1617]       seqsToCluster as lower-case ASCII (DNAbin 0x88=a 0x18=t 0x28=c 0x48=g)
  seq_names str    [6]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rda_reader import as_array, read_rda  # noqa: E402

REF = os.environ.get("DMPC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def mats(name):
    top = read_rda(f"{REF}/data/{name}.rda")[name]
    out = np.zeros((len(top.value), len(top.value[0].value), 4, 4))
    for i, per_idx in enumerate(top.value):
        for r, m in enumerate(per_idx.value):
            assert [x.value for x in m.attrs["dimnames"].value] == [list("atcg")] * 2
            out[i, r] = as_array(m)
    return out


def main():
    within, between = mats("withinClusTransProbs"), mats("betweenClusTransProbs")
    q = read_rda(f"{REF}/data/QmatrixPosada2001.rda")["QmatrixPosada2001"]
    assert [x.value for x in q.attrs["dimnames"].value] == [list("atcg")] * 2
    s = read_rda(f"{REF}/data/seqsToCluster.rda")["seqsToCluster"]
    code = np.zeros(256, np.uint8)
    for raw, ch in ((0x88, "a"), (0x18, "t"), (0x28, "c"), (0x48, "g")):
        code[raw] = ord(ch)
    seqs = np.stack([code[x.value] for x in s.value])
    assert seqs.min() > 0 and seqs.shape == (6, 1617)
    np.savez_compressed(os.path.join(HERE, "packaged_data.npz"), within=within, between=between,
                        Q=as_array(q), seqs=seqs, seq_names=np.array(s.attrs["names"].value))
    print("rows sum to 1:", np.abs(within.sum(-1) - 1).max(), np.abs(between.sum(-1) - 1).max())


if __name__ == "__main__":
    main()
