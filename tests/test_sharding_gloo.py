"""N > 1 host logic on CPU with the gloo backend (world_size 2): stream sharding, the one-time
weight-blob broadcast and result gathering -- the same code paths bench.py / model.to(broadcast_src=)
run over NCCL on GPUs."""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisper_medusa_b200 import _lib
from whisper_medusa_b200.parallel import broadcast_packed_weights, gather_stream_results, stream_ids_for_rank


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from whisper_medusa_b200.model import WhisperMedusaModel
        from whisper_medusa_b200.synthetic import preset_config, synthetic_state_dict
        from whisper_medusa_b200.weights import pack_blob

        cfg = preset_config("micro", heads=4)
        lib = _lib.load()
        h = C.c_void_p()
        wc = WhisperMedusaModel(cfg, None)._wm_config()
        assert lib.wm_create(C.byref(wc), -1, C.byref(h)) == 0      # layout-only handle (no GPU here)
        nbytes = lib.wm_weights_nbytes(h)
        blob = pack_blob(h, cfg, synthetic_state_dict(cfg, seed=4)) if rank == 0 else None   # only rank 0 reads the checkpoint
        got = broadcast_packed_weights(nbytes, 0, blob, torch.device("cpu"))
        lib.wm_destroy(h)
        n_streams = 7
        ids = stream_ids_for_rank(rank, world, n_streams)
        fake = [[sid, sid * 10] for sid in ids]                       # stands in for decoded token lists
        allr = gather_stream_results(fake, ids, n_streams)
        q.put((rank, int(got.to(torch.int64).sum()), ids, allr))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_weight_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, sum0, ids0, all0), (r1, sum1, ids1, all1) = res
    assert sum0 == sum1 and sum0 > 0                      # every rank holds the same blob
    assert sorted(ids0 + ids1) == list(range(7)) and not set(ids0) & set(ids1)
    assert all0 == [[s, s * 10] for s in range(7)] and all1 is None


def test_stream_partition_properties():
    for world in (1, 2, 4, 8):
        parts = [stream_ids_for_rank(r, world, 64) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
