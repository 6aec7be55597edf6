"""Re-hosted ``distributed_image_translation.py`` (reference :26-46,326-638): one process per GPU under
``torchrun`` (LOCAL_RANK/RANK/WORLD_SIZE from the environment), NCCL process group, rank-0 logging and saving,
``--load_*`` weight resume.  Instead of four DistributedDataParallel wrappers (which crash on the second forward,
SURVEY.md F4) the trainer all-reduces the flat gradient buffer of each stepped network."""
import os

import torch
import torch.distributed as dist

from ._cli import build_parser, run_training


def parse_args(argv=None):
    return build_parser("distributed").parse_args(argv)


def setup(rank, world_size):
    """reference :26-40 (MASTER_ADDR/PORT default to localhost:12355)."""
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12355")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device(f"cuda:{rank}"))


def cleanup():
    if dist.is_initialized():
        dist.destroy_process_group()


def main(argv=None):
    args = parse_args(argv)
    if "LOCAL_RANK" in os.environ:                     # reference :618-620
        args.distributed = True
        args.local_rank = int(os.environ["LOCAL_RANK"])
        args.world_size = int(os.environ.get("WORLD_SIZE", args.world_size))
    if args.distributed:
        setup(args.local_rank, args.world_size)
    tr = None
    try:
        tr = run_training(args, "distributed", rank=args.local_rank if args.distributed else 0,
                          world=args.world_size if args.distributed else 1)
        return tr
    finally:
        if args.distributed:
            if tr is not None:
                tr.close()           # captured graphs must go before the NCCL communicator
            dist.barrier()
            cleanup()


if __name__ == "__main__":
    main()
