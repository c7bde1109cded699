"""Host logic of the multi-GPU path on CPU: the stripe partition map and the gather/un-permute
arithmetic, exercised through a real 2-rank `gloo` job (the N > 1 path uses the same calls over NCCL)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import vpt_b200 as V
from vpt_b200.renderer import stripe_rows_of_rank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stripes_cover_every_row_exactly_once():
    for H, R, S in [(1080, 8, 8), (1080, 2, 16), (720, 4, 8), (17, 3, 4), (5, 8, 1)]:
        rows = [stripe_rows_of_rank(H, r, R, S) for r in range(R)]
        assert len({len(x) for x in rows}) == 1, "every rank holds the same (padded) row count"
        valid = np.concatenate([x[x < H] for x in rows])
        assert sorted(valid.tolist()) == list(range(H))


def test_single_rank_is_identity():
    assert np.array_equal(stripe_rows_of_rank(33, 0, 1, 16), np.arange(33))


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.environ["VPT_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    from vpt_b200.renderer import stripe_rows_of_rank
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, S = 24, 37, 4
    # stand-in for the per-rank render: value of pixel (x, y) is a pure function of the GLOBAL index, as in the kernels
    rows = stripe_rows_of_rank(H, rank, world, S)
    local = torch.zeros(len(rows) * W, 3)
    for lr, y in enumerate(rows):
        if y < H:
            idx = torch.arange(W) + y * W
            local[lr * W:(lr + 1) * W] = torch.stack([idx.float(), idx.float() * 0.5, idx.float() + 7], dim=1)
    gathered = torch.empty(world * local.shape[0], 3)
    dist.all_gather_into_tensor(gathered, local)                # the one collective of the path
    full = torch.full((H * W, 3), -1.0)
    n_local = local.shape[0]
    for r in range(world):                                       # host mirror of k_unpermute
        rr = stripe_rows_of_rank(H, r, world, S)
        for lr, y in enumerate(rr):
            if y < H:
                full[y * W:(y + 1) * W] = gathered[r * n_local + lr * W: r * n_local + (lr + 1) * W]
    idx = torch.arange(H * W).float()
    want = torch.stack([idx, idx * 0.5, idx + 7], dim=1)
    assert torch.equal(full, want), "gathered frame differs from the single-rank frame"
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


def test_two_rank_gloo_gather_reconstructs_the_frame(tmp_path):
    script = tmp_path / "worker.py"; script.write_text(WORKER)
    env = dict(os.environ, VPT_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2
