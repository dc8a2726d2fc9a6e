"""Stand-in for the third-party `aenum` package (absent in this image).

TEST INFRASTRUCTURE ONLY: lets tests/golden/make_golden.py import the unmodified
reference from /root/reference. The reference only needs stdlib-enum semantics plus
`extend_enum`, which is reached solely for user-defined object types
(multigrid/utils/enum.py:62, multigrid/core/world_object.py:57-58).
"""
from enum import *  # noqa: F401,F403
from enum import EnumMeta  # noqa: F401


def extend_enum(cls, name, value):
    raise NotImplementedError("aenum.extend_enum is not available in the golden-vector shim")
