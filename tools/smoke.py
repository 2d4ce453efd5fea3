#!/usr/bin/env python
"""Runs __graft_entry__.smoke() only (no rebuild): the command the compute-sanitizer runs wrap."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402

__graft_entry__.smoke()
