"""Host-side logic of the N > 1 path on CPU (world_size 2, gloo): slice sharding, the flat parameter / gradient buffers
and the bucketed gradient all-reduce (rpnet_b200.train.GradBuckets) — the only exchange step of the path.
No kernels run here; the GPU parity of the step itself is in tests/test_gpu_train.py."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(T=2):
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
                n_iter_refinement=T, soft_mask=False, mask_refinement_correlation_radius=5)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions_slices():
    from rpnet_b200.train import shard_range
    for total in (1, 7, 32, 96):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(96, 3, 8) == (36, 48)            # cfg5: 12 contiguous slices per GPU


def test_flat_params_layout_and_buckets():
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.train import BUCKET_GROUPS, FlatParams, GradBuckets, used_parameters
    torch.manual_seed(0)
    net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg())
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = FlatParams(net)
    flat.alias_grads()
    assert flat.numel >= 34_808_000 and flat.numel - 34_808_000 < 4 * len(flat.names)      # used parameters
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k]), k                                                  # values preserved
    for n, p in used_parameters(net):
        o, k = flat.offsets[n]
        assert p.data_ptr() == flat.param[o:o + k].data_ptr() and p.grad.data_ptr() == flat.grad[o:o + k].data_ptr()
    assert net.cre.w_context[0].weight.grad is None and net.cre.out[0].weight.grad is None     # D4
    # load_state_dict writes through into the flat buffer
    sd = {k: (v + 1 if v.is_floating_point() else v) for k, v in before.items()}
    net.load_state_dict(sd)
    o, k = flat.offsets['encoder.Conv1.conv.0.weight']
    assert torch.equal(flat.param[o:o + k], sd['encoder.Conv1.conv.0.weight'].reshape(-1))
    b = GradBuckets(flat, 1)
    assert len(b.ranges) == len(BUCKET_GROUPS)
    assert sum(hi - lo for lo, hi in b.ranges) == flat.numel


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.train import FlatParams, GradBuckets
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    torch.manual_seed(0)
    net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg())
    flat = FlatParams(net)
    flat.alias_grads()
    g = torch.Generator().manual_seed(100 + rank)
    flat.grad.copy_(torch.randn(flat.numel, generator=g))
    buckets = GradBuckets(flat, world)
    for i in range(len(buckets.ranges)):          # the order the backward finishes them
        buckets.ready(i)
    buckets.finish()
    want = sum(torch.randn(flat.numel, generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    ok = torch.allclose(flat.grad, want, rtol=1e-6, atol=1e-6)
    # every rank ends with the same gradient views on its parameters
    w = net.encoder.Conv3.conv[0].weight
    o, k = flat.offsets['encoder.Conv3.conv.0.weight']
    ok = ok and torch.equal(w.grad.reshape(-1), flat.grad[o:o + k])
    q.put((rank, bool(ok), float(flat.grad.double().sum())))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2]                   # identical reduced gradients on both ranks
