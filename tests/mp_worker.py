"""Worker of the one-process-per-GPU parity test (tests/test_gpu_parity.py::test_one_process_per_gpu_matches_oracle):
every process drives one GPU through reo_comm_init_rank (NCCL inside the library), K1 sharded over the ranks, and checks
the whole identify_degs output against the C oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(rank, world, tmpdir):
    os.environ["REO_K1_SHARD_MIN"] = "0"          # shard the staging even on these small inputs
    import __graft_entry__ as ge
    from conftest import small_case
    pkg = ge.load_package()
    oracle, co = ge.load_oracle()
    idfile = os.path.join(tmpdir, "nccl_id.bin")
    if rank == 0:
        with open(idfile + ".tmp", "wb") as f:
            f.write(pkg.nccl_unique_id())
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise RuntimeError("rank 0 never published the NCCL id")
            time.sleep(0.05)
    uid = open(idfile, "rb").read()
    h = pkg.Reo(rank, seed=pkg.synth.TIE_SEED)
    h.comm_init(rank, world, uid)

    def check(out, want):
        assert out.iters == want["iters"], (out.iters, want["iters"])
        assert np.array_equal(out.result[:, :, 2:11], want["result"][:, :, 2:11]), "tables differ"
        assert np.array_equal(out.updown, want["updown"]) and np.array_equal(out.final_ref, want["final_ref"])
        a, b = out.result[:, :, :2], want["result"][:, :, :2]
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) <= 1e-12

    # (genes, n1, n2, n3, reference mask): small one-sided, all genes (symmetric sweep), large subset (permuted panel),
    # three levels (one-vs-rest)
    cases = [(700, 40, 45, 0, lambda r: np.arange(r) % 5 == 0),
             (1700, 33, 31, 0, lambda r: np.ones(r, bool)),
             (2600, 20, 44, 0, lambda r: np.arange(r) % 9 != 0),
             (500, 12, 20, 17, lambda r: np.arange(r) % 3 == 0)]
    for ci, (r, n1, n2, n3, mk) in enumerate(cases):
        data, group = small_case(100 + ci, r, n1, n2, n3=n3)
        levels, gid = oracle.group_levels(group)
        gnum = len(levels)
        ref = mk(r)
        thr = co.thresholds_for(gid, gnum, 0.01)
        want = co.identify_degs(data, gid, gnum, thr, 1.0, 0.05, ref, 128, 5, seed=pkg.synth.TIE_SEED)
        check(h.identify_degs(data, gid, gnum, ref, 0.01, 1.0, 0.05, 128, 5), want)
        # stage-level: full build and incremental update against the oracle's tables
        h.stage(data, gid, gnum)
        m2 = np.arange(r) % 4 != 1
        tab, _ = co.block_tables(data, gid, gnum, thr, np.nonzero(m2)[0], seed=pkg.synth.TIE_SEED)
        assert np.array_equal(h.tables(0, m2, thresholds=thr), tab)
        assert np.array_equal(h.tables(0, ref, thresholds=thr, mask_to=m2), tab)
    h.close()
    with open(os.path.join(tmpdir, f"ok_{rank}"), "w") as f:
        f.write("ok")
