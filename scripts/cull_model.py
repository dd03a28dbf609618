"""Estimate the fraction of pairs culled by bounding-box tests for different block shapes
(rows x cols of Hilbert-consecutive atoms), N atoms uniform in a cubic box, cutoff L/2 - 0.1."""
import sys, numpy as np

def hilbert_keys(ix, iy, iz, bits):
    # Skilling's transpose -> Hilbert index
    X = [ix.astype(np.uint32).copy(), iy.astype(np.uint32).copy(), iz.astype(np.uint32).copy()]
    n = 3
    M = np.uint32(1 << (bits - 1))
    Q = M
    while Q > 1:
        P = np.uint32(Q - 1)
        for i in range(n):
            m = (X[i] & Q) != 0
            # invert
            X[0] = np.where(m, X[0] ^ P, X[0])
            # exchange
            t = (X[0] ^ X[i]) & P
            t = np.where(m, 0, t).astype(np.uint32)
            X[0] ^= t
            X[i] ^= t
        Q = np.uint32(Q >> 1)
    for i in range(1, n):
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[n - 1] & Q) != 0, t ^ np.uint32(Q - 1), t)
        Q = np.uint32(Q >> 1)
    for i in range(n):
        X[i] ^= t
    key = np.zeros(ix.shape, dtype=np.uint64)
    for b in range(bits - 1, -1, -1):
        for i in range(n):
            key = (key << np.uint64(1)) | ((X[i] >> np.uint32(b)) & np.uint32(1)).astype(np.uint64)
    return key

def boxes(pos, g):
    n = len(pos) // g
    p = pos[: n * g].reshape(n, g, 3)
    return p.min(1), p.max(1)

def main():
    N = int(sys.argv[1]); L = (N / 0.05) ** (1 / 3); cut = L / 2 - 0.1
    rng = np.random.default_rng(0)
    pos = rng.random((N, 3)) * L
    cells = np.minimum((pos / L * 128).astype(np.int64), 127)
    key = hilbert_keys(cells[:, 0], cells[:, 1], cells[:, 2], 7)
    pos = pos[np.argsort(key, kind="stable")]
    for rg, cg in ((32, 64), (32, 32), (32, 16), (16, 16), (32, 8)):
        rmin, rmax = boxes(pos, rg); cmin, cmax = boxes(pos, cg)
        # sample rows
        sel = rng.choice(len(rmin), size=min(400, len(rmin)), replace=False)
        far = 0; tot = 0
        for i in sel:
            dc = np.abs(0.5 * (rmin[i] + rmax[i]) - 0.5 * (cmin + cmax))
            h = 0.5 * (rmax[i] - rmin[i]) + 0.5 * (cmax - cmin)
            g = np.maximum(np.minimum(dc - h, L - (dc + h)), 0)
            far += np.count_nonzero((g * g).sum(1) > cut * cut); tot += len(cmin)
        print(N, f"rows {rg} x cols {cg}: culled {far / tot:.4f} of pairs (beyond cutoff: {1 - 4/3*np.pi*(cut/L)**3:.4f})", flush=True)
main()
