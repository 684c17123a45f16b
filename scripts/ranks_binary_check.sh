#!/bin/bash
# one-shot check of the drop-in binary's --ranks path after moving the fork behind the file read
B=ltr-lowrank-sdp_b200/lib/lorads_b200
L=gpurun_out/r2_fork_after_read.log
: > $L
run() { name=$1; shift; s=$(date +%s%N); timeout 20 $B "$@" > /tmp/o.txt 2> /tmp/e.txt; rc=$?; e=$(date +%s%N);
  echo "== $name rc=$rc wall_ms=$(( (e - s) / 1000000 )) :: $*" >> $L
  grep -E "Reading SDPA|1.Primal Objective|2.Dual Objective|Constraint Violation\(1\)|all_time:|OuterIter:.*InnerIter|^ADMM Iter|End Program|Time limit" /tmp/o.txt | tail -9 >> $L; head -5 /tmp/e.txt >> $L; }
G="--phase1Tol 1e-2 --heuristicFactor 10"
run G11_r1 tests/golden/instances/G11.dat-s $G
run G11_r2 tests/golden/instances/G11.dat-s $G --ranks 2
run ctl_r1 tests/golden/instances/control_like_12_6.dat-s
run ctl_r2 tests/golden/instances/control_like_12_6.dat-s --ranks 2
run del14_r1 bench_data/delaunay_n14.dat-s --phase1Tol 1e+1 --heuristicFactor 100 --timesLogRank 0.25
run del14_r2 bench_data/delaunay_n14.dat-s --phase1Tol 1e+1 --heuristicFactor 100 --timesLogRank 0.25 --ranks 2
cat $L
