"""Multi-GPU inside the boundary (fmpc_multi_*): one blocking call, one handle + host thread per device,
contiguous shards.  With explicit nu0 the result equals the single-device call bit for bit.  The sharding logic is also
exercised on a one-GPU box by opening two device slots on GPU 0."""
import numpy as np
import pytest

from cases import small_problem

pytestmark = pytest.mark.gpu


def make(pk, cls, c, nb, **kw):
    return cls(c["A1"], c["A2"], c["B"], c["Q"], c["R"], c["Qf"], c["u_min"], c["u_max"], c["T"], c["x_min"], c["x_max"],
               max_batch=nb, **kw)


def device_lists(pk):
    out = [[0, 0], [0, 0, 0]]                 # several slots on one GPU: shard logic without a second device
    nd = pk.device_count()
    if nd >= 2:
        out.append(list(range(min(nd, 8))))
    return out


def test_multi_step_equals_single_device_bit_for_bit(pk):
    c = small_problem(61, 28, 144, 20, 37, 0.5, warm=True, xf=True)      # 37 instances: uneven shards
    hb = make(pk, pk.FastMPCBatch, c, c["nb"])
    ref = hb.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], niters=4)
    hb.close()
    for devs in device_lists(pk):
        hm = make(pk, pk.FastMPCMulti, c, c["nb"], devices=devs)
        assert hm.ngpus == len(devs)
        out = hm.step(c["x0"], c["x0_pre"], c["w"], c["xf"], c["X0"], c["U0"], c["nu0"], niters=4)
        assert np.array_equal(out["U"], ref["U"]) and np.array_equal(out["X"], ref["X"])
        assert np.array_equal(out["iters"], ref["iters"]) and np.array_equal(out["status"], ref["status"])
        for use_nccl in (False, True):
            st = hm.stats(use_nccl=use_nccl)
            assert [s["n_solves"] for s in st] == [hm.shard(c["nb"], g)[1] for g in range(len(devs))]
            assert sum(s["n_solves"] for s in st) == c["nb"]
            assert sum(s["newton_iters"] for s in st) == int(ref["iters"].sum())
            assert sum(sum(s["status_hist"]) for s in st) == c["nb"]
        # a smaller batch than device slots: trailing shards are empty
        sub = hm.step(c["x0"][:2], c["x0_pre"][:2], c["w"][:2], c["xf"][:2], c["X0"][:2], c["U0"][:2], c["nu0"][:2], niters=4)
        assert np.array_equal(sub["U"], ref["U"][:2])
        hm.close()


def test_multi_resident_loop_equals_single_device(pk):
    c = small_problem(62, 12, 9, 7, 11, 0.5)
    rs = np.random.RandomState(5)
    hb = make(pk, pk.FastMPCBatch, c, c["nb"])
    hm = make(pk, pk.FastMPCMulti, c, c["nb"], devices=device_lists(pk)[-1])
    for k in range(4):
        x0 = 0.3 * rs.randn(c["nb"], c["n"])
        nu = rs.rand(c["nb"], c["T"] * c["n"])
        a = hb.step_resident(x0, nu0=nu, reset=(k == 0), full=True, niters=3)
        b = hm.step_resident(x0, nu0=nu, reset=(k == 0), full=True, niters=3)
        assert np.array_equal(a["u0"], b["u0"]) and np.array_equal(a["U"], b["U"]) and np.array_equal(a["iters"], b["iters"])
    # nu0 = NULL: one stream per device, device 0 = MATLAB's default stream => shard 0 equals the single-device result
    o1 = hb.step_resident(x0, reset=True, niters=3)
    o2 = hm.step_resident(x0, reset=True, niters=3)
    n0 = hm.shard(c["nb"], 0)[1]
    assert np.array_equal(o1["u0"][:n0], o2["u0"][:n0])
    hb.close(); hm.close()
