"""ncu target: ONE kernel family on 2^LOG2 elements.  usage: python tools/ncu_one.py <kernel> [log2=24] [reps=2]
kernels: interp interp_ring r2n n2r p2a p2a_ring p2v v2p neigh neigh_ring a2p"""
import subprocess
import sys

sys.argv = [sys.argv[0], sys.argv[2] if len(sys.argv) > 2 else "24", sys.argv[1], sys.argv[3] if len(sys.argv) > 3 else "2", "1", "default"]
exec(open("tools/ab_multi.py").read())
