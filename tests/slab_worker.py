"""torchrun worker for the multi-GPU parity test: every rank builds the same synthetic volume, the ranks solve it
together with medpy_b200.distributed over NCCL, rank 0 writes energy + gathered mask."""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from medpy_b200 import distributed as md, synthetic
    shape = tuple(int(s) for s in sys.argv[1].split("x"))
    case = sys.argv[2]
    out = sys.argv[3]
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    vol = synthetic.two_blob_volume(shape, seed=1, with_prob=(case == "regional"))
    kw = dict(image=vol["image"], boundary="difference_exponential", sigma=vol["sigma"])
    if case == "regional":
        kw.update(prob=vol["prob"], alpha=vol["alpha"])
    energy, mask = md.graphcut_slab(vol["fg"], vol["bg"], **kw)
    if dist.get_rank() == 0:
        numpy.savez(out, energy=energy, mask=mask)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
