"""Re-export of the packed containers (they live in the product package because the C-ABI
host layer needs them; the oracle only consumes them).  TEST INFRASTRUCTURE."""
from sbayes_b200.packed import *          # noqa: F401,F403
from sbayes_b200.packed import (Problem, ChainState, ChainConfig, GeoPriorSpec, NA, OP_NAMES, N_OPS)  # noqa: F401
