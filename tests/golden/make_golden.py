#!/usr/bin/env python3
"""Generate tests/golden/full_graph.npz from the reference's shipped fixtures.

Runs ONLY in the authoring container (needs /root/reference); the GPU box uses
the committed .npz.  Source data: /root/reference/ndt_feature/data/FULL GRAPH/
  mapping{k}.jff          LazyGrid maps written by upstream NDTMap::writeToJFF
  mapping{k}.T            node pose            (boost text archive, last 16 tokens = col-major 4x4)
  mapping{k}local_odom.T  odometry between node k and k+1
  mapping{k}local_fuse.T  the fuser's (matchFusion) estimate between node k and k+1
JFF layout as decoded in SURVEY.md Appendix B.  Only data is extracted (no reference source).
"""
import os, sys
import numpy as np

REF = "/root/reference/ndt_feature/data/FULL GRAPH"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_graph.npz")


def read_jff(path):
    d = open(path, "rb").read()
    assert d[:10] == b"#JFF V0.50", d[:10]
    assert np.frombuffer(d[10:14], "<i4")[0] == 3  # LazyGrid
    hdr = np.frombuffer(d[14:86], "<f8")  # size_m[3], cell[3], center[3]
    rec = np.frombuffer(d[566:], np.uint8)
    assert rec.size % 181 == 0
    rec = rec.reshape(-1, 181)
    has = rec[:, 136:140].copy().view("<i4").ravel() == 1
    tri = rec[:, 40:88].copy().view("<f8").reshape(-1, 6)
    mu = rec[:, 88:112].copy().view("<f8").reshape(-1, 3)
    n = rec[:, 128:132].copy().view("<i4").ravel()
    occ = rec[:, 152:156].copy().view("<f4").ravel()
    ctr = rec[:, 0:16].copy().view("<f4").reshape(-1, 4)[:, :3]
    return hdr, has, tri, mu, n, occ, ctr


def read_T(path):
    tok = open(path).read().split()
    return np.array(tok[-16:], float).reshape(4, 4).T


def main():
    out = {}
    for k in range(8):
        hdr, has, tri, mu, n, occ, ctr = read_jff(f"{REF}/mapping{k}.jff")
        out[f"hdr{k}"] = hdr
        idx = np.nonzero(has)[0].astype(np.int32)
        out[f"gidx{k}"] = idx              # linear record index (x-major, z-minor) of Gaussian cells
        out[f"mean{k}"] = mu[idx]
        out[f"cov{k}"] = tri[idx]          # xx,xy,xz,yy,yz,zz
        out[f"n{k}"] = n[idx]
        nz = np.nonzero(occ != 0)[0].astype(np.int32)
        out[f"occidx{k}"] = nz             # cells with non-zero log-odds occupancy
        out[f"occ{k}"] = occ[nz]
        out[f"occn{k}"] = n[nz]
        out[f"ctr{k}"] = ctr[idx]
        out[f"T{k}"] = read_T(f"{REF}/mapping{k}.T")
        out[f"Todom{k}"] = read_T(f"{REF}/mapping{k}local_odom.T")
        out[f"Tfuse{k}"] = read_T(f"{REF}/mapping{k}local_fuse.T")
        print(k, "gaussian cells", idx.size, "occ cells", nz.size, "hdr", hdr)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
