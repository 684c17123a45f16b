mkdir -p gpurun_out/e2e
for n in general_sparse_n60 multiblock_sdp multiblock_lp theta_n30; do
  ltr-lowrank-sdp_b200/lib/lorads_b200 tests/golden/instances/$n.dat-s > gpurun_out/e2e/$n.mine.log 2>&1
  OPENBLAS_NUM_THREADS=1 oracle/_ref/lorads_ref tests/golden/instances/$n.dat-s > gpurun_out/e2e/$n.ref.log 2>&1
done
