"""Entry point of one barrier-task worker process (the shim's counterpart of Spark's Python worker)."""
import sys

from .sql import run_barrier_task

if __name__ == "__main__":
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    run_barrier_task(rank, world, port, sys.argv[4], sys.argv[5])
