"""Batch driver: rational eigenvectors + eigenvalues for every squarefree level in a range, one
JSON file per level -- the command line of the reference's scripts/generate_eigenvalues.py
(169-231), with GPUs instead of a CPU process pool: run one process per GPU
(torchrun or RANK/WORLD_SIZE in the environment) and levels are dealt round-robin."""
import argparse
import logging
import os
import pathlib
import random
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

from ternary_birch_b200.batch import find_rational_eigenvectors, generate_jobs, is_squarefree  # noqa: E402

LOGGER = logging.getLogger(__name__)

if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s] [%(levelname)s] %(message)s",
                        datefmt="%Y-%m-%d %H:%M:%S")
    parser = argparse.ArgumentParser()
    parser.add_argument("savepath", type=pathlib.Path)
    parser.add_argument("--lower", type=int, default=500000, help="Lower bound of integers to check")
    parser.add_argument("--upper", type=int, default=1000000, help="Upper bound of integers to check")
    parser.add_argument("--num-primes", type=int, default=None)
    parser.add_argument("--upto", type=int, default=None,
                        help="Compute eigenvalues up to this bound. Defaults to the Sturm bound for the level.")
    parser.add_argument("--seed", type=int, default=random.randint(0, 2**32 - 1))
    args = parser.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if "LOCAL_RANK" in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", os.environ["LOCAL_RANK"])
    levels = [n for n in range(args.lower, args.upper) if is_squarefree(n, num_primes=args.num_primes)]
    mine = levels[rank::world]
    LOGGER.info("rank %s of %s: %s of %s levels", rank, world, len(mine), len(levels))
    args.savepath.mkdir(parents=True, exist_ok=True)
    for n, level in enumerate(mine):
        try:
            result = find_rational_eigenvectors(generate_jobs(level, seed=args.seed, upto=args.upto))
        except Exception:
            LOGGER.exception("Unexpected error occurred while processing level %s.", level)
        else:
            with open(args.savepath / f"{result.job.level}.json", "w") as f:
                f.write(result.json())
            LOGGER.info("(%s / %s) Finished computing eigenvalues at level %s", n + 1, len(mine), level)
