"""Imports the UNMODIFIED reference (read-only /root/reference/jorldy) from a writable copy.

Only used by make_golden.py in the build container; never at test/bench time on the GPU box
(where /root/reference does not exist).  A copy is needed because the reference's registries
write `_*_dict.txt` next to themselves at import (core/agent/__init__.py:24 etc.).
"""
import os
import shutil
import sys
import tempfile

REF_ROOT = "/root/reference/jorldy"


def import_reference():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference not present (golden vectors can only be minted in the build container)")
    dst = os.path.join(tempfile.gettempdir(), "jorldy_ref_copy")
    if not os.path.isdir(dst):
        shutil.copytree(REF_ROOT, dst, ignore=shutil.ignore_patterns("mlagents", "__pycache__"))
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import core.agent as agent_mod      # noqa: E402
    import core.buffer as buffer_mod    # noqa: E402
    import core.network as network_mod  # noqa: E402
    return agent_mod, buffer_mod, network_mod
