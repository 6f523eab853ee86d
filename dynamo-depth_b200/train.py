"""Entry point with the reference's command line (reference: train.py).  Single GPU:
    python train.py -d kitti -n run0 ...
one process per GPU on one node (torchrun sets LOCAL_RANK / LOCAL_WORLD_SIZE / WORLD_SIZE):
    torchrun --standalone --nproc-per-node 8 train.py -d waymo --cuda_ids 0 1 2 3 4 5 6 7
"""
import os

from torch.distributed import destroy_process_group, init_process_group

from options import DynamoOptions
from Trainer import Trainer


def ddp_setup():
    init_process_group(backend="nccl")


def ddp_cleanup():
    destroy_process_group()


if __name__ == "__main__":
    opt = DynamoOptions().parse()
    opt.local_world_size = int(os.environ.get("LOCAL_WORLD_SIZE", 1))
    opt.ddp = opt.local_world_size > 1
    if "LOCAL_RANK" in os.environ:          # torchrun passes the rank through the environment only
        opt.local_rank = int(os.environ["LOCAL_RANK"])
    assert len(opt.cuda_ids) == opt.local_world_size, \
        f"opt.cuda_ids(={opt.cuda_ids}) does not match opt.local_world_size(={opt.local_world_size})"
    if opt.ddp:
        ddp_setup()
    trainer = Trainer(opt)
    trainer.train()
    if opt.ddp:
        ddp_cleanup()
