"""N>1 host logic on CPU: world_size 2 over gloo.

Every rank builds the SAME cutout table from the same seed, takes its shard, computes its share of the loss
(scaled by the GLOBAL cutout count) and of the image gradient, and one all-reduce(sum) of each reproduces the
single-process result.  The per-rank compute here is the CPU oracle (the CUDA engine needs a GPU); the pieces under
test are the product's sharding (cutouts.shard_rows), its sampler determinism and its all-reduce helper.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import guidance as guidance_oracle
from oracle.make_golden import TINY, tiny_state_dict
from perceptor_b200 import cutouts
from perceptor_b200.guidance import _all_reduce_sum

N_CUT, H, W = 7, 48, 40


def _inputs():
    g = torch.Generator().manual_seed(0)
    images = torch.rand(2, 3, H, W, generator=g)
    targets = torch.nn.functional.normalize(torch.randn(2, TINY["embed"], generator=g))
    return images, targets, torch.tensor([1.0, 0.5])


def _shard_loss_and_grad(rows, n_total, images, targets, weights, sd):
    img = images.clone().requires_grad_()
    if len(rows) == 0:
        return torch.zeros(()), torch.zeros_like(images)
    enc = guidance_oracle.encode_cutouts(img, rows, sd, TINY["image_size"], TINY["patch"], TINY["layers"], TINY["heads"])
    dist_ = (enc[:, None] - targets[None, :]).norm(dim=2).div(2).arcsin().square().mul(2)
    loss = (dist_ * weights).sum() / (n_total * targets.shape[0])  # global mean, local sum
    loss.backward()
    return loss.detach(), img.grad


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    images, targets, weights = _inputs()
    sd = tiny_state_dict()
    rows = cutouts.sample_cutouts(torch.Generator().manual_seed(11), 2, H, W, N_CUT, 1.0, 16, 40)
    sl = cutouts.shard_rows(rows.shape[0], rank, world)
    loss, grad = _shard_loss_and_grad(rows[sl].tolist(), rows.shape[0], images, targets, weights, sd)
    loss = loss.reshape(1).clone()
    _all_reduce_sum(loss, dist.group.WORLD)
    _all_reduce_sum(grad, dist.group.WORLD)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss.numpy(), grad=grad.numpy(), rows=rows,
             start=sl.start, stop=sl.stop)
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["rows"], r1["rows"]), "ranks must build identical cutout tables"
    assert (int(r0["start"]), int(r0["stop"]), int(r1["start"]), int(r1["stop"])) == (0, 7, 7, 14)
    assert np.array_equal(r0["loss"], r1["loss"]) and np.array_equal(r0["grad"], r1["grad"])
    images, targets, weights = _inputs()
    rows = r0["rows"]
    loss, grad = _shard_loss_and_grad(rows.tolist(), rows.shape[0], images, targets, weights, tiny_state_dict())
    assert abs(float(r0["loss"][0]) - float(loss)) <= 1e-6 * abs(float(loss))
    assert float(np.abs(r0["grad"] - grad.numpy()).max()) <= 1e-6 * float(grad.abs().max()) + 1e-9


def _worker_images(rank, world, port, out_dir):
    """shard="images": the rank holds ONLY its own image; the table is drawn for the concatenated batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    images, targets, weights = _inputs()
    mine = images[rank:rank + 1]
    sd = tiny_state_dict()
    rows = cutouts.sample_cutouts(torch.Generator().manual_seed(11), world, H, W, 6, 1.0, 16, 40)
    local = cutouts.local_rows(rows, rank, world, b_offset=rank)
    cutouts.validate_rows(local, 1, H, W)  # every local row addresses the one local image
    loss, grad = _shard_loss_and_grad(local.tolist(), rows.shape[0], mine, targets, weights, sd)
    loss = loss.reshape(1).clone()
    _all_reduce_sum(loss, dist.group.WORLD)  # the only collective of the step
    np.savez(os.path.join(out_dir, f"img_rank{rank}.npz"), loss=loss.numpy(), grad=grad.numpy(), rows=rows)
    dist.destroy_process_group()


def test_two_rank_image_sharding_needs_no_gradient_collective(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_images, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "img_rank0.npz"), np.load(tmp_path / "img_rank1.npz")
    assert np.array_equal(r0["rows"], r1["rows"]) and np.array_equal(r0["loss"], r1["loss"])
    images, targets, weights = _inputs()
    rows = r0["rows"]
    loss, grad = _shard_loss_and_grad(rows.tolist(), rows.shape[0], images, targets, weights, tiny_state_dict())
    assert abs(float(r0["loss"][0]) - float(loss)) <= 1e-6 * abs(float(loss))
    both = np.concatenate([r0["grad"], r1["grad"]])  # rank r's gradient IS the gradient of image r
    assert float(np.abs(both - grad.numpy()).max()) <= 1e-6 * float(grad.abs().max()) + 1e-9
