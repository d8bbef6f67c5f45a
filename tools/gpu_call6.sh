set -x
O=gpurun_out/r02f
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "interp or config4 or healpy or smoke" > $O/pytest_interp.log 2>&1; echo "pytest rc=$?" >> $O/pytest_interp.log
V=variants
timeout 300 python tools/ab_multi.py 28 interp 5 3 default $V/lib_interp_rows.so > $O/ab_interp.log 2>&1
timeout 300 python tools/ab_multi.py 28 interp_ring 5 3 default $V/lib_interp_rows.so > $O/ab_interp_ring.log 2>&1
ls -la $O
