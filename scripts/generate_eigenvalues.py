"""Batch run: rational eigenvectors and their eigenvalues for every squarefree level of a range, one
``<level>.json`` per level -- the command line and the documents of the reference's
scripts/generate_eigenvalues.py (169-231), on GPUs: start one process per GPU (torchrun, or RANK /
WORLD_SIZE / LOCAL_RANK in the environment) and the levels are dealt round-robin to the ranks."""
import argparse
import logging
import os
import pathlib
import random
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

from ternary_birch_b200 import batch  # noqa: E402

LOG = logging.getLogger("generate_eigenvalues")


def main() -> int:
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s] [%(levelname)s] %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("savepath", type=pathlib.Path)
    ap.add_argument("--lower", type=int, default=500000, help="Lower bound of integers to check")
    ap.add_argument("--upper", type=int, default=1000000, help="Upper bound of integers to check")
    ap.add_argument("--num-primes", type=int, default=None)
    ap.add_argument("--upto", type=int, default=None,
                    help="Compute eigenvalues up to this bound. Defaults to the Sturm bound for the level.")
    ap.add_argument("--seed", type=int, default=random.randint(0, 2**32 - 1))
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if "LOCAL_RANK" in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", os.environ["LOCAL_RANK"])
    levels = [n for n in range(args.lower, args.upper) if batch.is_squarefree(n, num_primes=args.num_primes)]
    mine = levels[rank::world]
    LOG.info("rank %s of %s: %s of %s levels", rank, world, len(mine), len(levels))
    args.savepath.mkdir(parents=True, exist_ok=True)
    failures = 0
    for done, level in enumerate(mine, 1):
        try:
            doc = batch.run_level(batch.plan_level(level, seed=args.seed, upto=args.upto))
        except Exception:
            failures += 1
            LOG.exception("level %s failed", level)
            continue
        batch.write_document(doc, args.savepath)
        LOG.info("(%s / %s) level %s: %s rational eigenvectors, eigenvalues up to %s", done, len(mine), level,
                 doc["rational_eigenvectors_count"], doc["job"]["eigenvalues_upto"])
    return 0 if failures == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
