"""CPU test of the persistent GEMM's work decomposition (csrc/separable.cu: SkSched / SkIter, exported as
tg_gemm_schedule): every (tile, accumulation chunk) is covered exactly once, stream-K pieces of a tile have
distinct slots 0..nparts-1, and the CTAs' loads are balanced.  No GPU needed: the same host/device code
enumerates the units in the kernel's producer, MMA and epilogue roles."""
import ctypes as C

import numpy as np
import pytest

from temgymcore_b200 import _lib as L

SHAPES = [
    # (M, N, K, f16)                       what it is
    (1024, 2048, 20000, 1),              # BASELINE C2: 128 tiles on 148 SMs -> all tiles stream-K
    (128, 2048, 20000, 1),               # C2 row shard of one rank at N = 8: 16 tiles
    (256, 2048, 20000, 1),               # N = 4 shard / e2e row block
    (2048, 4096, 32768, 1),              # C3 batch: 512 tiles = 3 waves + 68
    (2048, 4096, 3392, 1),               # C3 last batch
    (100, 72, 40, 1), (128, 128, 128, 1), (130, 260, 1000, 1), (1, 1, 1, 1), (4096, 128, 64, 1),
    (1024, 2048, 20000, 0), (300, 500, 4000, 0), (128, 2048, 777, 0),
    (1024, 1024, 30720, 1), (512, 1024, 30720, 1), (512, 2048, 20000, 1),   # 3-product C2 and its row blocks: pieces + tails
    (19 * 128, 128, 5000, 1),            # 19 tiles
    (148 * 128, 128, 4096, 1),           # exactly one wave: no stream-K
    (149 * 128, 128, 4096, 1),           # one wave + 1 tile
]


def schedule(M, N, K, f16, sms=148, mode=2):
    lib = L.load()
    sched = (C.c_int32 * 10)()
    n = lib.tg_gemm_schedule(M, N, K, f16, sms, mode, None, 0, sched)
    assert n > 0
    units = np.zeros((n, 6), dtype=np.int32)
    n2 = lib.tg_gemm_schedule(M, N, K, f16, sms, mode, units.ctypes.data, n, sched)
    assert n2 == n
    keys = ("tiles_n", "T", "nkb", "nch", "G", "R", "q", "Tl", "qh", "maxparts")
    return units, dict(zip(keys, list(sched)))


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("M,N,K,f16", SHAPES)
def test_every_chunk_covered_once(M, N, K, f16, mode):
    units, s = schedule(M, N, K, f16, mode=mode)
    T, nch = s["T"], s["nch"]
    assert T == -(-M // 128) * -(-N // 128)
    cover = np.zeros((T, nch), dtype=np.int32)
    slots = {}
    for cta, tile, c0, c1, slot, nparts in units:
        assert 0 <= cta < s["G"] <= 148
        assert 0 <= tile < T and 0 <= c0 < c1 <= nch, (tile, c0, c1)
        cover[tile, c0:c1] += 1
        slots.setdefault(tile, []).append((slot, nparts))
        assert nparts <= s["maxparts"]
        if tile >= s["R"]:
            assert (c0, c1, slot, nparts) == (0, nch, 0, 1)       # data-parallel tile: whole K, direct store
    assert (cover == 1).all()
    for tile, sl in slots.items():
        nparts = sl[0][1]
        assert all(n == nparts for _, n in sl)
        assert sorted(x for x, _ in sl) == list(range(nparts)), (tile, sl)   # the counter reaches nparts exactly


@pytest.mark.parametrize("M,N,K,f16", [s for s in SHAPES if s[2] >= 4000])
def test_load_balance(M, N, K, f16):
    units, s = schedule(M, N, K, f16)
    load = np.zeros(s["G"], dtype=np.int64)
    for cta, tile, c0, c1, slot, nparts in units:
        load[cta] += c1 - c0
    total = s["T"] * s["nch"]
    assert load.sum() == total
    ideal = total / s["G"]
    # the busiest CTA sets the time: within 3 % (+ one chunk) of a perfect split over the CTAs used
    assert load.max() <= 1.03 * ideal + 1, (load.max(), ideal, s)
    if total >= 148 * 8:
        assert s["G"] >= 126          # plain split-K uses floor(148 / T) * T CTAs, the other schedules all 148


def test_c2_uses_every_sm_and_heads_walk_in_lockstep():
    units, s = schedule(1024, 2048, 20000, 1)
    assert s["G"] == 148 and s["R"] == 128 and s["maxparts"] <= 3
    heads = units[units[:, 4] == 0]
    assert len(heads) == 128 and (heads[:, 2] == 0).all() and (heads[:, 3] == s["q"]).all()


def test_default_policy_splits_only_small_shapes():
    """mode 1 (the library's default): the full C2 / C3 images keep whole tiles (86 % of a wave: measured faster
    unsplit), row shards and row blocks are split over all SMs."""
    for M, expect_split in ((1024, False), (512, True), (256, True), (128, True)):
        _, s = schedule(M, 2048, 20000, 1, mode=1)
        assert (s["R"] > 0) == expect_split, (M, s)
        if expect_split:
            assert s["G"] >= 128 and s["R"] == s["T"] and s["Tl"] == 0      # plain split-K: T * floor(148 / T) CTAs
    _, s = schedule(2048, 4096, 32768, 1, mode=1)
    assert s["R"] == 0 and s["G"] == 148
    _, s = schedule(1024, 2048, 20000, 1, mode=0)
    assert s["R"] == 0 and s["G"] == 128


def test_plain_split_k_groups_walk_k_in_lockstep():
    """Few tiles: CTA c owns k-range c // T of tile c % T, so the T CTAs of a group cover the same chunks."""
    units, s = schedule(256, 2048, 20000, 1, mode=1)          # a 256-row block: 32 tiles, 4 ranges
    assert s["T"] == 32 and s["maxparts"] == 4 and s["G"] == 128
    for cta, tile, c0, c1, slot, nparts in units:
        assert tile == cta % 32 and slot == cta // 32 and nparts == 4
        assert c0 == slot * s["q"] and c1 == min(s["nch"], c0 + s["q"])
    _, s3 = schedule(256, 2048, 20000, 1, mode=3)             # mode 3: head / tail instead
    assert s3["G"] == 148 and s3["Tl"] > 0


def test_split_k_with_tails_fills_the_idle_sms():
    """Opt-in mode 4 (measured slower than the plain split at the power cap, kept for other boards).
    The 3-product C2 GEMM (64 tiles): 2 full pieces per tile on 128 CTAs walking k in lockstep, and the last chunks
    of every tile -- one common k-range -- laid end to end over the other 20 SMs; every CTA within a chunk of ideal."""
    kch = L.load().tg_gemm_chunk_k()
    K3 = 3 * kch * -(-10000 // kch)
    units, s = schedule(1024, 1024, K3, 1, mode=4)
    assert s["T"] == 64 and s["G"] == 148 and s["R"] == 64 and s["Tl"] > 0
    q, nch = s["q"], s["nch"]
    pieces = units[units[:, 0] < 128]
    assert len(pieces) == 128
    for cta, tile, c0, c1, slot, nparts in pieces:
        assert tile == cta % 64 and slot == cta // 64 and (c0, c1) == (slot * q, slot * q + q)
    tails = units[units[:, 0] >= 128]
    assert (tails[:, 2] >= 2 * q).all() and (tails[:, 3] <= nch).all() and (tails[:, 4] >= 2).all()
    load = np.zeros(148, dtype=np.int64)
    for cta, tile, c0, c1, slot, nparts in units:
        load[cta] += c1 - c0
    assert load.max() <= -(-64 * nch // 148)
    _, s4 = schedule(1024, 1024, K3, 1, mode=1)
    assert s4["G"] == 128 and s4["Tl"] == 0 and s4["q"] > q


def test_streamk_can_be_disabled_by_shape():
    # shallow K: not worth splitting
    units, s = schedule(128, 2048, 512, 1)
    assert s["R"] == 0 and s["G"] == 16 and len(units) == 16


@pytest.mark.parametrize("M,N,K,rtm", [(1024, 1024, 30720, 2), (1024, 1024, 30720, 1), (1024, 2048, 20000, 2),
                                       (900, 1024, 30720, 2), (1024, 2048, 20000, 1), (640, 1000, 9000, 2),
                                       (2048, 4096, 32768, 2)])
def test_row_block_streaming_rounds(M, N, K, rtm):
    """mode 100 + r (the host pipeline's single-launch streaming): rounds of r tile rows, every round split along K
    over the machine; every (tile, chunk) once, slots complete, and a CTA meets its tiles in increasing round order
    with the same k-range in every round (its column-factor slice stays in L2)."""
    units, s = schedule(M, N, K, 1, mode=100 + rtm)
    tiles_m, tiles_n = -(-M // 128), s["tiles_n"]
    T, nch = s["T"], s["nch"]
    Tr = tiles_n * rtm
    if tiles_m <= rtm or Tr > 148:
        pytest.skip("not a streaming shape")
    assert s["R"] == T and s["G"] == Tr * s["maxparts"] <= 148
    cover = np.zeros((T, nch), dtype=np.int32)
    last_round = {}
    krange = {}
    slots = {}
    for cta, tile, c0, c1, slot, nparts in units:
        cover[tile, c0:c1] += 1
        rnd = tile // Tr
        assert tile % Tr == cta % Tr and slot == cta // Tr and nparts == s["maxparts"]
        assert last_round.get(cta, -1) < rnd
        last_round[cta] = rnd
        assert krange.setdefault(cta, (c0, c1)) == (c0, c1)
        slots.setdefault(tile, []).append(slot)
    assert (cover == 1).all()
    for tile, sl in slots.items():
        assert sorted(sl) == list(range(s["maxparts"]))
    if M == 1024 and N == 1024 and rtm == 2:
        assert s["maxparts"] == 9 and s["G"] == 144         # 4 rounds of 16 complex tiles x 9 k-ranges


# ---- ragged schedule of the tile-binned sum (TG_METHOD_TENSOR_BINNED): tiles own different numbers of chunks --------
def ragged(chunks, sms=148):
    lib = L.load()
    ch = np.ascontiguousarray(chunks, dtype=np.int32)
    n = lib.tg_gemm_schedule_ragged(len(ch), ch.ctypes.data, sms, None, 0, None, 0)
    assert n >= 0
    units = np.zeros((max(n, 1), 8), dtype=np.int32)
    readers = np.full(max(n, 1) * (sms + 1), -1, dtype=np.int32)
    n2 = lib.tg_gemm_schedule_ragged(len(ch), ch.ctypes.data, sms, units.ctypes.data, n, readers.ctypes.data,
                                     len(readers))
    assert n2 == n
    return units[:n], readers


RAGGED_CASES = [
    ("c3_like", lambda r: r.integers(8, 20, size=512)),
    ("mostly_empty", lambda r: np.where(r.random(512) < 0.9, 0, r.integers(1, 40, size=512))),
    ("one_huge_tile", lambda r: np.array([0, 0, 5000, 0, 3, 0])),
    ("fewer_chunks_than_ctas", lambda r: np.array([1, 0, 2, 0, 0, 1, 3])),
    ("single_chunk", lambda r: np.array([0, 1, 0])),
    ("uniform", lambda r: np.full(148, 7)),
    ("ramp", lambda r: np.arange(300) % 17),
]


@pytest.mark.parametrize("sms", [148, 132, 7])
@pytest.mark.parametrize("name,gen", RAGGED_CASES)
def test_ragged_schedule_covers_every_chunk_once(name, gen, sms):
    chunks = np.asarray(gen(np.random.default_rng(7)), dtype=np.int64)
    P = np.concatenate([[0], np.cumsum(chunks)])
    tot = int(P[-1])
    units, readers = ragged(chunks, sms)
    cover = np.zeros(tot, dtype=np.int32)
    load = np.zeros(sms, dtype=np.int64)
    by_tile, written = {}, {}
    ri = 0
    for cta, tile, c0, c1, slot, nparts, wslot, first in units:
        assert 0 <= cta < sms and P[tile] <= c0 < c1 <= P[tile + 1], (cta, tile, c0, c1)
        cover[c0:c1] += 1
        load[cta] += c1 - c0
        assert 0 <= slot < nparts and cta == first + slot
        by_tile.setdefault(tile, []).append((slot, nparts))
        if nparts > 1:
            assert wslot in (2 * cta, 2 * cta + 1)
            assert wslot not in written, "two partial units of one launch share a scratch slot"
            written[wslot] = (tile, slot)
        rd = readers[ri:ri + nparts]
        ri += nparts
        by_tile[tile][-1] += (tuple(rd),)
    assert (cover == 1).all()
    q = max(1, -(-tot // sms))
    assert load.max() <= q                                   # even split of the chunk axis
    for tile, parts in by_tile.items():
        nparts = parts[0][1]
        assert sorted(p[0] for p in parts) == list(range(nparts))          # the arrival counter reaches nparts
        if nparts > 1:
            for slot, _, rd in parts:
                # whoever arrives last reads, part by part, exactly the slots the parts were parked in
                assert [written[r] for r in rd] == [(tile, s) for s in range(nparts)]
    assert set(np.nonzero(chunks)[0]) == set(by_tile)        # empty tiles get no unit
