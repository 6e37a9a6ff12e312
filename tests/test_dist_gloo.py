"""World-size-2 gloo tests (CPU) of the N > 1 host logic: the flat-buffer gradient reducer over the ranges the native
backward completes part by part, and DistributedDataParallel construction around the drop-in module."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fcmae_oracle as fo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build_model():
    import mmearth_train_b200 as m
    args = fo.make_args(None, "uncertainty")
    return m.convnextv2_atto(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True, patch_size=8,
                             img_size=56, args=args, loss_fn=m.UncertaintyWeightingStrategy(12))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mmearth_train_b200.dist import FlatGradReducer, seed_scale
        model = _build_model()
        plan = model._plan(4)
        ranges = plan.backward_ranges()
        n = model._n_flat
        # ranges tile the flat buffer in reverse layer order: tail (heads/decoder/proj) first, patch embedding last
        assert ranges[0][1] == n and ranges[2][0] == 0 and ranges[0][0] == ranges[1][1] and ranges[1][0] == ranges[2][1]
        names = {nm: off for nm, _s, off, _d in plan.params()}
        assert ranges[0][0] == names["proj.weight"] and ranges[1][0] == names["encoder.stages.2.0.dwconv.kernel"]
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(n, generator=g) * seed_scale()          # each rank's local gradient, pre-scaled by 1/world
        local = flat.clone()
        red = FlatGradReducer(ranges, n)
        for part in range(3):                                       # reverse layer order, asynchronous
            red.reduce_part(flat, part)
        red.wait()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(flat, sum(gathered), atol=1e-6)       # == mean of the unscaled gradients (DDP semantics)
        assert red.bytes_reduced == n * 4
        with pytest.raises(ValueError):
            FlatGradReducer([(0, 10), (12, n)], n)
        # DDP (main_pretrain.py:306-310) must accept the module and manage only the token parameter
        ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=False)
        managed = [nm for nm, p in ddp.module.named_parameters() if nm not in ddp.parameters_to_ignore]
        assert managed == ["_ddp_token"], managed
        assert ddp.module is model
        out[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        out[rank] = "".join(traceback.format_exception(type(e), e, e.__traceback__))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_reducer_and_ddp_wrap_world2(native_lib):
    world = 2
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("gloo worker hung")
    assert dict(out) == {0: "ok", 1: "ok"}, dict(out)


def _loop_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from argparse import Namespace
        from mmearth_train_b200 import engine
        from oracle import make_engine_golden as meg
        from tests.test_engine_golden import _CpuScaler, _Replay
        torch.set_num_threads(2)
        cfg = dict(model="convnextv2_atto", img_size=56, patch_size=8, out_modalities=["sentinel2"], loss_aggr="unweighted")
        orc = fo.build_oracle(**cfg)
        fo.init_like_reference(orc, seed=3)                                       # same weights on every rank
        n_iter, B = 3, 2
        batches = [fo.synthetic_batch(B, 56, ["sentinel2"], seed=40 + 10 * rank + i) for i in range(n_iter)]   # rank-local data
        g = torch.Generator().manual_seed(7 + rank)                               # rank-local masks (main_pretrain.py:202-203)
        noises = [torch.randn(B, 49, generator=g) for _ in range(n_iter)]
        model = torch.nn.parallel.DistributedDataParallel(_Replay(orc, noises))   # main_pretrain.py:306-310
        args = Namespace(update_freq=1, lr=3e-4, min_lr=1e-6, warmup_epochs=1, epochs=4, mask_ratio=0.6, no_ffcv=True)
        opt = torch.optim.AdamW(meg.param_groups_weight_decay(orc, 0.05), lr=args.lr, betas=(0.9, 0.95))
        writer = meg._Writer() if rank == 0 else None        # main_pretrain.py:255-260: the log writer exists on rank 0 only
        stats, loss_dict, _, _ = engine.train_one_epoch(model, None, [(i, b) for i, b in enumerate(batches)], opt,
                                                        torch.device("cpu"), 1, False, _CpuScaler(), log_writer=writer,
                                                        args=args, lag=2, quiet=True)
        # the meters are summed over ranks at the end (helpers.py:37-49): every rank reports the global mean ...
        both = [None, None]
        rows = [r["train_loss"] for r in writer.rows if r["head"] == "loss"] if writer is not None else None
        dist.all_gather_object(both, (stats["loss"], rows, float(sum(p.detach().double().sum() for p in orc.parameters()))))
        assert abs(both[0][0] - both[1][0]) < 1e-12, both
        # ... the per-iteration loss rank 0 logs is the mean over ranks (helpers.all_reduce_mean, called by every rank
        # every step whether or not it holds a writer) ...
        assert both[1][1] is None and len(both[0][1]) == n_iter
        assert abs(sum(both[0][1]) / n_iter - both[0][0]) < 1e-6
        # ... and DDP's averaged gradients keep the replicas identical
        assert abs(both[0][2] - both[1][2]) < 1e-9
        m = engine.SmoothedValue()
        m.update(float(rank + 1))
        m.synchronize_between_processes()
        assert m.count == 2 and m.total == 3.0
        out[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        out[rank] = "".join(traceback.format_exception(type(e), e, e.__traceback__))
    finally:
        dist.destroy_process_group()


def test_training_loop_world2_meters_and_logged_losses():
    """engine.train_one_epoch under torch.distributed (gloo, world size 2) around a DDP-wrapped CPU oracle."""
    world = 2
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_loop_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("gloo worker hung")
    assert dict(out) == {0: "ok", 1: "ok"}, dict(out)
