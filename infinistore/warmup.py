"""Alias of infinistore_b200.warmup (see infinistore/__init__.py)."""
from infinistore_b200.warmup import *  # noqa: F401,F403
from infinistore_b200.warmup import main  # noqa: F401

if __name__ == "__main__":
    main()
