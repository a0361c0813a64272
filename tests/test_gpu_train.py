"""-m gpu: the reference-facing call, CaffeNet.train(solver_index, FloatBlob[])
(CaffeNet.java:120, CaffeNetTest.testTrain :270-321): host blobs in, one
Solver::Step, with the PyTorch gradient producer standing in for
Net::ForwardBackward.  Checks the step against the oracle given the SAME
gradient, and that training actually reduces the loss."""
import numpy as np
import pytest
import torch

from gpu_util import assert_bits_equal, to_host

pytestmark = pytest.mark.gpu


def test_train_lenet_from_prototxt(cos, oracle, tmp_path):
    from caffeonspark_b200 import harness, nets
    solver = nets.write_prototxts("lenet", str(tmp_path))
    net = cos.CaffeNet(solver, "", "", 1, 1, 0, True, cos.CaffeNet.NONE, -1, 0)  # allocate(prototxt...) path
    try:
        assert net.param_count() == nets.EXPECTED_PARAM_COUNT["lenet"]
        assert net.connect(net.localAddresses())
        assert net.init(0, True)
        net.set_option("zero_diff", 0)  # keep the gradient so the test can feed it to the oracle
        prod = harness.make_producer("lenet", net)
        desc = net.desc
        rng = np.random.RandomState(0)
        x = torch.from_numpy(rng.rand(64, 1, 28, 28).astype(np.float32)).pin_memory()
        y = torch.from_numpy(rng.randint(0, 10, (64, 1, 1, 1)).astype(np.float32)).pin_memory()
        losses = []
        for t in range(20):
            w_before, h_before = to_host(net.data()), to_host(net.history())
            net.diff().zero_()
            torch.cuda.synchronize()
            rate = net.learning_rate()
            assert net.train(0, [x, y]), net.last_error()
            assert net.synchronize(), net.last_error()  # train() is pipelined: it returns once the batch left host memory
            losses.append(net.last_loss())
            g = to_host(net.diff())  # the gradient the producer accumulated (diff_ not cleared)
            oracle.apply_update(0, desc.param_count, w_before, g.copy(), h_before, desc.counts, desc.lr_mult,
                                desc.decay_mult, np.float32(rate), np.float32(desc.momentum),
                                np.float32(desc.weight_decay))
            assert_bits_equal(to_host(net.data()), w_before, f"weights after train() iter {t}")
            assert_bits_equal(to_host(net.history()), h_before, f"history after train() iter {t}")
        assert net.iter() == 20
        assert np.isfinite(losses).all() and losses[-1] < losses[0] and losses[-1] < 50.0  # CaffeNetTest: loss < 50
        assert prod is not None
    finally:
        net.deallocate()


def test_train_two_ranks_in_process(cos, oracle):
    """Two executors, real train() calls from two threads (CaffeProcessor runs
    one solver thread per device), gradients from the PyTorch producer."""
    import concurrent.futures as cf
    from caffeonspark_b200 import harness, nets
    desc = nets.solver_desc("cifar10_quick")
    netz = [cos.CaffeNet(desc, "", "", 1, 2, r, True, cos.CaffeNet.SOCKET, -1, 0) for r in range(2)]
    try:
        for n in netz:
            n.set_option("grid", 8)
            n.set_option("block", 128)
            n.set_option("barrier_timeout_ms", 20000)
            n.set_option("initial_gather", 0)
        prods = [harness.make_producer("cifar10_quick", n, seed=5, use_graph=False) for n in netz]  # same init on both ranks
        table = [n.localAddresses() for n in netz]
        with cf.ThreadPoolExecutor(2) as ex:
            assert all(ex.map(lambda r: netz[r].connect([table[p][r] if p != r else "" for p in range(2)]), range(2)))
        for n in netz:
            assert n.all_gather_weights(0), n.last_error()
        for n in netz:
            assert n.synchronize(), n.last_error()
        rng = np.random.RandomState(1)
        batches = [(torch.from_numpy(rng.rand(100, 3, 32, 32).astype(np.float32)),
                    torch.from_numpy(rng.randint(0, 10, (100,)).astype(np.float32))) for _ in range(2)]

        def run(r):
            out = []
            for t in range(6):
                assert netz[r].train(0, list(batches[r])), netz[r].last_error()
                if t == 0:  # pipelined train(): the loss arrives later; wait for the first one explicitly
                    assert netz[r].synchronize(), netz[r].last_error()
                out.append(netz[r].last_loss())  # otherwise: the newest loss that has ARRIVED (may lag a step)
            assert netz[r].synchronize(), netz[r].last_error()
            out.append(netz[r].last_loss())
            return out

        with cf.ThreadPoolExecutor(2) as ex:
            losses = list(ex.map(run, range(2)))
        w0, w1 = to_host(netz[0].data()), to_host(netz[1].data())
        assert_bits_equal(w0, w1, "ranks hold identical weights after synchronous SGD")
        assert all(np.isfinite(l).all() for l in losses)
        assert losses[0][-1] < losses[0][0]
        assert prods
    finally:
        with cf.ThreadPoolExecutor(2) as ex:
            list(ex.map(lambda n: n.deallocate(), netz))


def test_train_pipelined_equals_synchronous(cos, oracle):
    """Row f3: train() double-buffers the input staging (H2D of batch t+1 on a copy stream while step t
    computes) and returns once its batch has left host memory.  The caller may then overwrite the host blobs
    (the Scala side recycles them, CaffeProcessor.scala:442-452): do exactly that, and require the run to be
    bit-identical to the fully synchronous one (train_pipeline=0) on the same batch sequence."""
    from caffeonspark_b200 import harness, nets
    desc = nets.solver_desc("lenet")
    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True  # the two runs must agree bit for bit: no atomics in the producer
    rng = np.random.RandomState(3)
    xs = [rng.rand(64, 1, 28, 28).astype(np.float32) for _ in range(8)]
    ys = [rng.randint(0, 10, (64, 1, 1, 1)).astype(np.float32) for _ in range(8)]
    results = []
    for pipeline in (0, 1):
        net = cos.CaffeNet(desc)
        try:
            assert net.connect(net.localAddresses())
            net.set_option("train_pipeline", pipeline)
            harness.make_producer("lenet", net, seed=11)
            x, y = torch.empty(64, 1, 28, 28).pin_memory(), torch.empty(64, 1, 1, 1).pin_memory()
            losses = []
            for t in range(8):
                x.copy_(torch.from_numpy(xs[t]))  # ONE pair of host blobs, overwritten right after train() returns
                y.copy_(torch.from_numpy(ys[t]))
                assert net.train(0, [x, y]), net.last_error()
                x.fill_(float("nan"))             # a late H2D would poison the run
            assert net.synchronize(), net.last_error()
            losses.append(net.last_loss())
            assert net.iter() == 8
            results.append((to_host(net.data()), to_host(net.history()), losses[-1]))
        finally:
            net.deallocate()
    torch.backends.cudnn.deterministic = det
    assert np.isfinite(results[0][0]).all()
    assert_bits_equal(results[1][0], results[0][0], "pipelined vs synchronous weights")
    assert_bits_equal(results[1][1], results[0][1], "pipelined vs synchronous history")
    assert results[0][2] == results[1][2]
