"""Alias of infinistore_b200.benchmark (see infinistore/__init__.py)."""
from infinistore_b200.benchmark import *  # noqa: F401,F403
from infinistore_b200.benchmark import main  # noqa: F401

if __name__ == "__main__":
    main()
