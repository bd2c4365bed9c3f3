"""Worker of tests/test_multigpu.py, one rank per GPU under torchrun:
shard a batch over the ranks, run the forward on every GPU, all-gather the fp32 logits with the library's own
NVLink collective (parallel.LogitsAllGather / eqxv_allgather_push) and compare them BITWISE with the single-GPU
result of the same images computed on rank 0."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import eqxvision_b200 as eb
    from eqxvision_b200 import parallel
    from tools import synthetic as syn

    per_rank = int(os.environ.get("EQXV_TEST_PER_RANK", "64"))
    arch = os.environ.get("EQXV_TEST_ARCH", "vit_base")
    total = per_rank * world
    if arch == "vit_base":
        sd = syn.vit_state_dict(embed_dim=768, depth=12, heads=12, num_classes=1000, seed=3)
    else:
        sd = syn.torchvision_state_dict(arch, seed=1)
    with tempfile.NamedTemporaryFile(suffix=".pth") as f:
        torch.save(sd, f.name)
        net = eb.models.vit_base(num_classes=1000, torch_weights=f.name) if arch == "vit_base" else \
            getattr(eb.models, arch)(torch_weights=f.name)
    net = eb.tree_inference(net, True)
    fwd = eb.filter_jit(eb.vmap(net, axis_name="batch"))
    keys = eb.random.split(eb.random.PRNGKey(0), per_rank)
    images = syn.synthetic_images(total, seed=31)                 # the same 512 images on every rank

    gathered = parallel.data_parallel_forward(lambda x: fwd(x, key=keys), images, gather=True)
    assert gathered.shape == (total, 1000) and gathered.is_cuda
    # a second and third round through the same communicator (buffer parity, flag epochs)
    again = parallel.data_parallel_forward(lambda x: fwd(x, key=keys), images, gather=True)
    third = parallel.all_gather_rows(fwd(parallel.shard(images), key=keys), total)
    assert torch.equal(again, gathered) and torch.equal(third, gathered)
    # uneven shards (padding path of all_gather_rows): 5 rows per rank but the last rank one fewer
    n_odd = 5 * world - 1
    lo, hi = parallel.shard_bounds(n_odd, rank, world)
    odd = parallel.all_gather_rows(gathered[lo:hi].contiguous(), n_odd)
    assert torch.equal(odd, gathered[:n_odd])

    # every rank holds the same gathered matrix
    digest = gathered.double().sum().reshape(1)
    digests = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(digests, digest)
    assert all(torch.equal(d, digests[0]) for d in digests)

    if rank == 0:
        single = torch.cat([fwd(images[i:i + per_rank], key=keys) for i in range(0, total, per_rank)])
        same = torch.equal(single, gathered)
        print(f"MULTIGPU world={world} arch={arch} images={total} bitwise_equal={same} "
              f"max_abs_diff={(single - gathered).abs().max().item():.3e}", flush=True)
        assert same
    dist.barrier()
    for ag, _ in parallel._gatherers.values():
        ag.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
