"""nuts333_b200 -- the NUTS 3.3.3 message path (write_user / write_room[_except] /
write_level colour rendering and fan-out, contains_swearing, site_banned /
user_banned) as hand-written sm_100a CUDA kernels behind a C-ABI.

    from nuts333_b200 import Context, Talker

The CUDA library is built in-tree by `nuts333_b200.build.build()`; there is no CPU path.
"""
from .api import (Context, Talker, Streams, NutsbError, pack,  # noqa: F401
                  OP_USER, OP_ROOM, OP_LEVEL, OF_FORCE_LISTEN, OF_SHOUT, OF_ABOVE, OF_GATE_IF_SET, OF_PAGER, OF_PLAIN,
                  UF_COLOUR, UF_LOGIN, UF_IGNALL, UF_IGNSHOUT, SAY, SHOUT, SEMOTE,
                  NEW, USER, WIZ, ARCH, GOD)
