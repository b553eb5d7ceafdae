"""CPU model of the tile-marking test of the tile-binned tensor-core sum (csrc/separable.cu: bin_mark_kernel): the
footprint of a beamlet is the axis-aligned ellipse {envelope >= threshold} (+1 px slack), and a 128-row x 64-column
tile is marked when the closest point of its pixel rectangle to the ellipse's centre lies inside the ellipse.  The
property the sum relies on: EVERY pixel whose envelope reaches the threshold lies in a marked tile (nothing above the
culling threshold is dropped), for row blocks too; and the marking is tight (a marked tile is never farther than the
slack from the footprint).  Same formulas, fp32 where the kernel uses fp32."""
import numpy as np
import pytest

BM, BIN_TN = 128, 64


def mark_tiles(e, thr_bits, H, W, row0, nrows):
    """numpy restatement of bin_mark_kernel for ONE beamlet: e = (e0..e5) of E(c, r) = e0 + e1 c + e2 r + e3 c^2 +
    e4 c r + e5 r^2 [bits]; thr_bits = the threshold on E.  Returns the set of (tm, tn) marked inside rows
    [row0, row0 + nrows)."""
    tiles_m, tiles_n = (nrows + BM - 1) // BM, (W + BIN_TN - 1) // BIN_TN
    tn_lo, tn_hi, tm_lo, tm_hi = 0, tiles_n - 1, 0, tiles_m - 1
    cx = cy = np.float32(0.0)
    hx = hy = np.float32(1e30)
    det = e[3] * e[5] - 0.25 * e[4] * e[4]
    if e[3] < 0.0 and e[5] < 0.0 and det > 0.0:
        cs = (0.5 * e[4] * e[2] - e[5] * e[1]) / (2.0 * det)
        rs = (0.5 * e[4] * e[1] - e[3] * e[2]) / (2.0 * det)
        d = (e[0] + 0.5 * (e[1] * cs + e[2] * rs)) - thr_bits
        if d < 0.0:
            return set()
        if np.isfinite(d) and np.isfinite(cs) and np.isfinite(rs):
            slack = (abs(cs) + abs(rs)) * 1.2e-7
            hc = np.sqrt(d * (-e[5]) / det) + 1.0 + slack
            hr = np.sqrt(d * (-e[3]) / det) + 1.0 + slack
            c_lo, c_hi = max(np.floor(cs - hc), 0.0), min(np.ceil(cs + hc), W - 1.0)
            r_lo, r_hi = max(np.floor(rs - hr), float(row0)), min(np.ceil(rs + hr), row0 + nrows - 1.0)
            if c_hi < c_lo or r_hi < r_lo:
                return set()
            tn_lo, tn_hi = int(c_lo) // BIN_TN, int(c_hi) // BIN_TN
            tm_lo, tm_hi = (int(r_lo) - row0) // BM, (int(r_hi) - row0) // BM
            cx, cy, hx, hy = np.float32(cs), np.float32(rs), np.float32(min(hc, 1e30)), np.float32(min(hr, 1e30))
    out = set()
    with np.errstate(over="ignore"):
        hh = hx * hy
        for tm in range(tm_lo, tm_hi + 1):
            r0 = np.float32(row0 + tm * BM)
            r1 = np.float32(row0 + min(tm * BM + BM - 1, nrows - 1))
            dy = max(max(r0 - cy, cy - r1), np.float32(0.0)) * hx
            for tn in range(tn_lo, tn_hi + 1):
                c0 = np.float32(tn * BIN_TN)
                c1 = np.float32(min(tn * BIN_TN + BIN_TN - 1, W - 1))
                dx = max(max(c0 - cx, cx - c1), np.float32(0.0)) * hy
                if dx * dx + dy * dy <= hh * hh * np.float32(1.000001):
                    out.add((tm, tn))
    return out


def random_envelope(rng, H, W):
    """a separable (no cross term) concave envelope with its vertex on or around the detector"""
    sc, sr = rng.uniform(3.0, 90.0, 2)                 # 1-bit half-widths in pixels
    cs, rs = rng.uniform(-0.2 * W, 1.2 * W), rng.uniform(-0.2 * H, 1.2 * H)
    peak = rng.uniform(-30.0, 0.0)
    e3, e5 = -1.0 / sc ** 2, -1.0 / sr ** 2
    return np.array([peak + e3 * cs * cs + e5 * rs * rs, -2.0 * e3 * cs, -2.0 * e5 * rs, e3, 0.0, e5])


@pytest.mark.parametrize("H,W,row0,nrows", [(512, 512, 0, 512), (300, 416, 0, 300), (512, 384, 96, 130), (700, 200, 256, 444)])
def test_every_pixel_above_the_threshold_lies_in_a_marked_tile(H, W, row0, nrows):
    rng = np.random.default_rng(11)
    cols, rows = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(row0, row0 + nrows, dtype=np.float64))
    n_marked = n_needed = 0
    for _ in range(60):
        e = random_envelope(rng, H, W)
        thr = rng.uniform(-60.0, -20.0)
        E = e[0] + e[1] * cols + e[2] * rows + e[3] * cols * cols + e[5] * rows * rows
        above = E >= thr
        need = {(int(r) // BM, int(c) // BIN_TN) for r, c in zip(*np.nonzero(above))}
        got = mark_tiles(e, thr, H, W, row0, nrows)
        assert need <= got, (e, thr, sorted(need - got))
        n_marked += len(got)
        n_needed += len(need)
        # tight: a marked tile holds a pixel within ~1.5 px (the slack) of the footprint
        ring = E >= thr - 2.0 * max(abs(e[3]), abs(e[5])) ** 0.5 * np.sqrt(max(e[0] - thr, 1.0)) * 4.0 - 4.0 * (abs(e[3]) + abs(e[5]))
        near = {(int(r) // BM, int(c) // BIN_TN) for r, c in zip(*np.nonzero(ring))}
        assert got <= near | need, (e, thr, sorted(got - near))
    assert n_needed > 0 and n_marked <= 1.35 * n_needed + 60      # not marking the bounding box's empty corners wholesale


def test_degenerate_envelopes_mark_everything_or_nothing():
    H, W = 256, 320
    allt = {(tm, tn) for tm in range(2) for tn in range(5)}
    # non-concave, or not finite: the beamlet is carried everywhere (NaN beamlets must poison the image)
    assert mark_tiles(np.array([0.0, 0.0, 0.0, 1e-3, 0.0, -1e-3]), -40.0, H, W, 0, H) == allt
    assert mark_tiles(np.array([np.nan] * 6), -40.0, H, W, 0, H) == allt
    # below the threshold everywhere, or off the detector
    assert mark_tiles(np.array([-100.0, 0.0, 0.0, -1e-2, 0.0, -1e-2]), -40.0, H, W, 0, H) == set()
    far = random_envelope(np.random.default_rng(0), H, W)
    far = np.array([far[0], far[1] + 2 * far[3] * 5000.0, far[2], far[3], 0.0, far[5]])   # vertex shifted by -5000 columns
    far[0] = -5.0 + far[1] ** 2 / (4 * -far[3]) * -1.0 + far[2] ** 2 / (4 * -far[5]) * -1.0
    assert mark_tiles(far, -40.0, H, W, 0, H) == set()
