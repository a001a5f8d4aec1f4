"""Multi-GPU pieces of the path that need two devices in one process (-m gpu; skipped on a one-GPU box):
the grouped NCCL all-reduce of the run totals (sk_allreduce_totals, what the `fasta` binary calls at the end
of a run) and shard linearity across devices.  The `fasta` binary's own multi-GPU run is covered by
tests/test_gpu_cli.py::test_all_visible_gpus_give_the_single_gpu_bytes (all visible GPUs vs SK_GPUS=1)."""
import ctypes as C

import pytest

import fuzzgen as G

pytestmark = pytest.mark.gpu


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_devices() < 2, reason="needs two GPUs")
def test_run_totals_are_merged_over_two_gpus_with_one_allreduce():
    import numpy as np
    from seqkit_b200 import Engine, _lib as L
    S = 96
    sheet, bcs = G.make_sheet(3, S, 8)
    engs = [Engine(device=d, max_stream_bytes=32 << 20, max_records=1 << 16, max_samples=S, aux_streams=False) for d in (0, 1)]
    try:
        lib = engs[0].lib
        want = np.zeros(S + 2, dtype=np.uint64)
        opts = L.DemuxOpts(-1, 0, 0, 0, 0)
        files = []
        for d, eng in enumerate(engs):
            eng.set_sheet(bcs)
            for k in range(2):  # contiguous pair ranges: device d owns [40000 d, 40000 (d + 1))
                first = 40000 * d + 20000 * k
                eng.synth(0, 20000, seed=4, first_pair=first, mate=1, with_bc=True)
                eng.synth(1, 20000, seed=4, first_pair=first, mate=2, with_bc=True)
                assert lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
                eng.wait()
                c = np.zeros(S + 2, dtype=np.uint64)
                assert lib.sk_download_counts(eng.ctx, 0, c.ctypes.data) == 0
                want += c
                assert lib.sk_counts_accumulate(eng.ctx, 0) == 0
        ctxs = (C.c_void_p * 2)(engs[0].ctx, engs[1].ctx)
        assert lib.sk_allreduce_totals(ctxs, 2) == 0, lib.sk_last_error(engs[0].ctx)
        for eng in engs:  # every context holds the run's counters
            got = np.zeros(S + 2, dtype=np.uint64)
            assert lib.sk_download_totals(eng.ctx, got.ctypes.data) == 0
            assert np.array_equal(got, want) and int(got[S]) == 80000
        # the same ranges on one device give the same counters (shard linearity across devices)
        eng = engs[0]
        assert lib.sk_totals_reset(eng.ctx) == 0
        for first in (0, 20000, 40000, 60000):
            eng.synth(0, 20000, seed=4, first_pair=first, mate=1, with_bc=True)
            eng.synth(1, 20000, seed=4, first_pair=first, mate=2, with_bc=True)
            assert lib.sk_demultiplex(eng.ctx, 0, C.byref(opts)) == 0
            eng.wait()
            assert lib.sk_counts_accumulate(eng.ctx, 0) == 0
        got = np.zeros(S + 2, dtype=np.uint64)
        assert lib.sk_download_totals(eng.ctx, got.ctypes.data) == 0
        assert np.array_equal(got, want)
    finally:
        for e in engs:
            e.close()
