set -x
O=gpurun_out/r02q
mkdir -p $O
V=variants
for ns in 1048576 262144 65536 4096; do
export HPGB_AB_NSIDE=$ns
for k in p2a p2v; do
  timeout 300 python tools/ab_multi.py 27 $k 5 3 default $V/lib_p2_notab.so $V/lib_p2t_map.so $V/lib_p2t_map_cg.so $V/lib_p2t_map_na.so $V/lib_p2t_h1s2.so $V/lib_p2t_h1.so $V/lib_p2t_h1_cg.so $V/lib_p2t_h1_na.so > $O/ab_${ns}_$k.log 2>&1; grep -v "result ==" $O/ab_${ns}_$k.log
done
done
