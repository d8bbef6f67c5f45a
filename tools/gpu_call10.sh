set -x
O=gpurun_out/r02k
mkdir -p $O
V=variants
timeout 300 python tools/ab_multi.py 28 interp 5 3 default $V/lib_interp_lut2.so $V/lib_interp_lut1.so $V/lib_interp_nolut.so > $O/ab_interp2.log 2>&1
timeout 300 python tools/ab_multi.py 28 interp_ring 5 3 default $V/lib_interp_lut2.so $V/lib_interp_lut1.so $V/lib_interp_nolut.so > $O/ab_interp_ring2.log 2>&1
cat $O/ab_interp2.log $O/ab_interp_ring2.log
