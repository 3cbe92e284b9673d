"""Phase timestamps of the tensor-core decoder kernels.  Debug build first:
    WB_LIB_NAME=libwispb200_timing.so WB_EXTRA_NVCC_FLAGS=-DWB_TC_TIMING python kaolin-wisp_b200/build.py

Runs one bench step in-process, then reads the clock64() stamps CTA 0 took in every tc_round
(entry, after __syncthreads, after issue, after the mbarrier wait) and prints the per-phase cycle averages per round slot.
Never used by the product or the tests."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["WISPB200_LIB"] = os.path.join(ROOT, "kaolin-wisp_b200", "lib", "libwispb200_timing.so")
sys.path.insert(0, ROOT)
sys.argv = ["bench.py", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"]
import bench
bench.run_ours(bench.parse())
import wisp_b200 as W
lib = W._cabi.lib()
N = 1024
buf = np.zeros((2, 2, N), np.int64)
rc = lib.wb_tc_timing_dump(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
for k, name, rounds in ((0, "fwd (per-tile extra barrier for the feature save not stamped)", 5), (1, "bwd", 10)):
    for t in (0, 1):
        ts = buf[k, t]
        n = int((ts > 0).sum())
        ts = ts[:n]
        per_tile = rounds * 4
        ntile = n // per_tile
        if ntile < 3:
            print(name, "thread", t, "too few stamps", n); continue
        a = ts[: ntile * per_tile].reshape(ntile, rounds, 4)[1:]          # skip the first tile
        sync = (a[:, :, 1] - a[:, :, 0]).mean(0)
        issue = (a[:, :, 2] - a[:, :, 1]).mean(0)
        wait = (a[:, :, 3] - a[:, :, 2]).mean(0)
        nxt = np.concatenate([a[:, 1:, 0], np.concatenate([a[1:, :1, 0], a[-1:, -1:, 3]], 0)], 1)
        epi = (nxt - a[:, :, 3])[:-1].mean(0)
        tile = (a[1:, 0, 0] - a[:-1, 0, 0]).mean()
        print(f"{name} stamp slot {t}: tiles={ntile} cycles/tile={tile:.0f}")
        if k == 1 and t == 1:     # phase of group 1 relative to group 0 (same SM clock): start of round 0 of tile j
            a0 = buf[1, 0][: ntile * per_tile].reshape(ntile, rounds, 4)
            a1 = buf[1, 1][: ntile * per_tile].reshape(ntile, rounds, 4)
            print("   group1 - group0 tile start (cycles):", [int(a1[j, 0, 0] - a0[j, 0, 0]) for j in range(0, ntile, max(1, ntile // 8))])
        for r in range(rounds):
            print(f"   round {r}: sync {sync[r]:7.0f}  issue {issue[r]:7.0f}  wait {wait[r]:7.0f}  epilogue(+next tile load after last) {epi[r]:7.0f}")

buf2 = np.zeros((2, N), np.int64)
if hasattr(lib, "wb_tc_timing_dump2") and lib.wb_tc_timing_dump2(buf2.ctypes.data_as(ctypes.c_void_p)) == 0:
    for k, name, rounds in ((0, "fwd", 5), (1, "bwd", 10)):
        t2 = buf2[k]; n = int((t2 > 0).sum()) // 3 * 3
        t0 = buf[k, 0]
        nt = min(n // (3 * rounds), int((t0 > 0).sum()) // (4 * rounds))
        if nt < 3:
            continue
        a = t0[: nt * rounds * 4].reshape(nt, rounds, 4)[1:]
        b = t2[: nt * rounds * 3].reshape(nt, rounds, 3)[1:]
        print(f"{name} thread 0, inside the issue phase (cycles): barrier->elected+fenced | UMMAs issued | commit | ->end of phase")
        for r in range(rounds):
            print("   round %d: %6.0f %6.0f %6.0f %6.0f" % (r, (b[:, r, 0] - a[:, r, 1]).mean(), (b[:, r, 1] - b[:, r, 0]).mean(),
                                                         (b[:, r, 2] - b[:, r, 1]).mean(), (a[:, r, 2] - b[:, r, 2]).mean()))

buf3 = np.zeros((2, N), np.int64)
if hasattr(lib, "wb_tc_timing_dump3") and lib.wb_tc_timing_dump3(buf3.ctypes.data_as(ctypes.c_void_p)) == 0:
    for k, name in ((0, "fwd"), (1, "bwd")):
        t3 = buf3[k]; n = int((t3 > 0).sum()) // 3 * 3
        if n < 30:
            continue
        a = t3[:n].reshape(-1, 3)[6:]
        print(f"{name} thread 0, hidden-layer epilogue (32 columns): tcgen05.ld+wait {(a[:, 1] - a[:, 0]).mean():.0f}  convert+store {(a[:, 2] - a[:, 1]).mean():.0f} cycles")

buf4 = np.zeros((2, 64, 16), np.int64)
if hasattr(lib, "wb_tc_timing_dump4") and lib.wb_tc_timing_dump4(buf4.ctypes.data_as(ctypes.c_void_p)) == 0:
    for k, name, rounds in ((0, "fwd", 5), (1, "bwd", 10)):
        print(f"{name} thread 0: cycles from chain start to each UTCHMMA issued (rounds of the 3rd tile)")
        for r in range(2 * rounds, 3 * rounds):
            row = buf4[k, r]; n = int((row > 0).sum())
            print("   round %2d:" % (r - 2 * rounds), [int(row[i] - row[0]) for i in range(1, n)])
