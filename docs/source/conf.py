"""Sphinx configuration (reference counterpart: docs/source/conf.py:21-28, which also mocks
the native module and torch so that autodoc runs without a GPU or a build)."""
import os
import sys

sys.path.insert(0, os.path.abspath("../.."))

project = "infinistore-b200"
author = "infinistore-b200 developers"
release = "0.1.0"

extensions = ["sphinx.ext.autodoc", "sphinx.ext.napoleon", "sphinx.ext.viewcode"]
autodoc_mock_imports = ["infinistore_b200._infinistore", "torch", "numpy", "fastapi", "uvicorn"]
autodoc_member_order = "bysource"
templates_path = []
exclude_patterns = []
html_theme = "alabaster"
