"""smoke() on the freshly built library, then kb_pool_trim: the pool's cached scratch goes back to the driver."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import __graft_entry__ as g
from kliff_b200 import ops

g.smoke()
free0 = torch.cuda.mem_get_info()[0]
ops.pool_trim(0)
print("pool_trim ok, free bytes", free0, "->", torch.cuda.mem_get_info()[0])
