"""
Peer-mailbox collectives (-m gpu): the scalar reductions of a target-sharded run as single-warp exchange kernels over
peer-mapped memory (k_peer_exchange, include/spectralkernels_b200.h "PEER MEMORY"), replacing the NCCL all-reduces behind
src/quadrature.jl:258-260 (max |I2-I1|) and src/adaptive.jl:183-198 (stopping index of the convergence scan).

Ranks as separate launches must not share a GPU (they wait for one another), so on this one-GPU tier the protocol runs
with the ranks emulated as the blocks of ONE cooperative launch (sk_comm_peer_selftest), and a real mailbox is exercised
with world size 1 (export, self-attach, every collective point of a kernel_values run goes through k_peer_exchange).
The 2/8-GPU run over real NVLink peers is tests/multi_gpu_check.py (profiles/r2_*_multi_gpu_check_*).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sk():
    import spectralkernels_jl_b200 as sk
    return sk


@pytest.mark.parametrize("nranks", [1, 2, 5, 8, 16])
def test_exchange_protocol_emulated_ranks(sk, nranks):
    eng = sk.Session(0)
    rng = np.random.default_rng(nranks)
    vals = rng.uniform(0, 1, nranks)
    maxbits = vals.view(np.uint64).copy()
    dist = rng.uniform(0, 1, nranks)
    lo = 1000
    top = rng.integers(lo - 1, lo + 5000, nranks)          # lo - 1: nothing unconverged on that rank
    top[0] = lo - 1
    rounds = 4
    out = eng.comm_peer_selftest(maxbits, dist.view(np.uint64), top, lo, rounds=rounds)
    want_max = int(maxbits.max()) + (rounds - 1)            # the hook adds (round - 1) to every rank's word
    live = top >= lo
    want_r = int(dist.view(np.uint64)[live].max()) if live.any() else 0
    want_n = int(np.maximum(top - lo + 1, 0).sum())
    for r in range(nranks):                                 # every rank holds the same reduced words
        assert int(out[r, 0]) == want_max
        assert int(out[r, 1]) == 0
        assert int(out[r, 2]) == want_r
        assert int(out[r, 3]) == want_n
        assert int(out[r, 4]) & 0xff == 0 and int(out[r, 4]) >> 8 == rounds     # no timeout, not void, last epoch seen
    if nranks > 1:
        # one rank's last exchange belongs to a chained launch that skipped itself: EVERY rank must see the exchange void
        # (mark in the epoch word) and the all-ones word in place of max |I2-I1| (fails every guard's "<" test)
        out = eng.comm_peer_selftest(maxbits, dist.view(np.uint64), top, lo, rounds=rounds, skip_rank=nranks - 1)
        for r in range(nranks):
            assert int(out[r, 0]) == 0xFFFFFFFFFFFFFFFF
            assert (int(out[r, 4]) >> 1) & 1 == 1 and int(out[r, 4]) & 1 == 0 and int(out[r, 4]) >> 8 == rounds


def test_world_size_one_mailbox_run_is_bitwise_identical(sk):
    """export + self-attach: sk_subinterval / sk_converge_scan / the early range reduction all take the peer path"""
    from spectralkernels_jl_b200.sharded import LibComm
    xs = np.random.default_rng(3).uniform(0, 1, 200_000)
    S = sk.Matern(1 / (np.pi / 2), 1.0, 1.5)
    cfg0 = sk.AdaptiveKernelConfig(S)
    t0 = []
    v0, e0 = sk.kernel_values(cfg0, xs, k0=1.0, trace=t0)
    cfg1 = sk.AdaptiveKernelConfig(S)
    eng = cfg1.engine
    comm = LibComm(eng, 0, 1, peer_handles=[eng.comm_peer_export()])
    assert comm.mode == "peer"
    assert comm.gather([1.5, 2.5, 3.5]) == [[1.5, 2.5, 3.5]]
    assert comm.max([4.0, -1.0]) == [4.0, -1.0] and comm.sum([0.25] * 9) == [0.25] * 9
    t1 = []
    v1, e1 = sk.kernel_values(cfg1, xs, k0=1.0, comm=comm, trace=t1)
    comm.close()
    assert np.array_equal(v0, v1) and np.array_equal(e0, e1)
    key = lambda tr: [(t["kind"], t["a"], t["b"], t.get("accepted"), t.get("hi_after")) for t in tr]
    assert key(t0) == key(t1)
    # after close the context is a plain single-GPU context again
    v2, e2 = sk.kernel_values(cfg1, xs, k0=1.0)
    assert np.array_equal(v0, v2)
