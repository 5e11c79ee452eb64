#!/bin/bash
# One gpurun session: tests, microbenchmark, bench variants, ncu captures.  Usage: tools/gpu_session.sh <tag> [steps...]
# Everything lands under gpurun_out/<tag>/ ; each step has its own timeout so that a hang cannot eat the box.
TAG=${1:-s}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
STEPS=${@:-"tests l2 variants bench ref ncu"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
for S in $STEPS; do
  case $S in
    tests)  timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt ;;
    tests_all) timeout 1800 python -m pytest tests -m gpu -q --timeout=600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/rc.txt ;;
    smoke)  timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/rc.txt ;;
    l2)     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_gather_bench tools/l2_gather_bench.cu 2> $OUT/l2_gather.err \
              && timeout 120 /tmp/l2_gather_bench > $OUT/l2_gather.json 2>> $OUT/l2_gather.err; echo "l2 rc=$?" >> $OUT/rc.txt ;;
    variants)
      # A/B of kernel variants through environment switches (bench.py --no-extras: cfg 3 only)
      i=0
      for CFG in "KGE_SPLIT_VARIANT=2" "KGE_SPLIT_VARIANT=4" "KGE_SPLIT_VARIANT=1"; do
        i=$((i+1))
        env $CFG timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity > $OUT/bench_ab$i.json 2> $OUT/bench_ab$i.err
        echo "ab$i [$CFG] rc=$?" >> $OUT/rc.txt
      done ;;
    ncu_final)
      B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity --eval-queries 512"
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
        --log-file $OUT/launches.csv $B > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?" >> $OUT/rc.txt
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"row_kernel_split|entity_kernel" -c 4 -o $OUT/prof_train \
        $B > $OUT/ncu_train.log 2>&1; echo "ncu train rc=$?" >> $OUT/rc.txt
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"count_ranks_kernel|rescore" -c 3 -o $OUT/prof_eval \
        $B > $OUT/ncu_eval.log 2>&1; echo "ncu eval rc=$?" >> $OUT/rc.txt
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_count_kernel|split_tf32" -c 3 -o $OUT/prof_gemm \
        python bench.py --workload complex_wn18rr --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity --eval-queries 2048 \
        > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" >> $OUT/rc.txt ;;
    mgpu)   # needs gpurun --gpus N
      NG=$(nvidia-smi -L | wc -l)
      timeout 1200 python -m pytest tests/test_multi_gpu.py -q -x --timeout=900 > $OUT/pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$?" >> $OUT/rc.txt
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $NG --no-cpu-baseline > $OUT/bench_g$NG.json 2> $OUT/bench_g$NG.err; echo "bench $NG gpus rc=$?" >> $OUT/rc.txt
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --impl reference --gpus $NG --steps 3 --warmup 1 > $OUT/bench_ref_g$NG.json 2> $OUT/bench_ref_g$NG.err; echo "ref $NG gpus rc=$?" >> $OUT/rc.txt ;;
    escale)  # does the row kernel speed up when the entity table fits the L2 comfortably?
      for NE in 3500 7000 14951 30000; do
        timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none \
          -k regex:"row_kernel_split|entity_kernel" -c 8 --csv --log-file $OUT/escale_$NE.csv \
          python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity --eval-queries 64 --entities $NE > $OUT/escale_$NE.log 2>&1
        echo "escale $NE rc=$?" >> $OUT/rc.txt
      done ;;
    phases) timeout 300 python tools/row_phases.py > $OUT/row_phases.txt 2>&1; timeout 300 python tools/row_phases.py --negatives 64 >> $OUT/row_phases.txt 2>&1; echo "phases rc=$?" >> $OUT/rc.txt ;;
    gemmbench) timeout 400 python bench.py --workload complex_wn18rr --no-extras --no-cpu-baseline --no-parity --eval-queries 4096 > $OUT/bench_complex.json 2> $OUT/bench_complex.err; echo "gemmbench rc=$?" >> $OUT/rc.txt
               timeout 600 python -m pytest tests -m gpu -q -k "tcgen05 or tensor_core or full_size_eval or runpy_wn18rr" --timeout=600 > $OUT/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?" >> $OUT/rc.txt ;;
    nscale)
      for NN in 64 128 512; do
        timeout 300 python bench.py --no-extras --no-cpu-baseline --no-parity --negatives $NN --eval-queries 256 > $OUT/bench_n$NN.json 2> $OUT/bench_n$NN.err
        echo "nscale $NN rc=$?" >> $OUT/rc.txt
      done ;;
    bench)  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/rc.txt ;;
    hostprof) timeout 300 python tools/profile_host.py > $OUT/hostprof.txt 2>&1; echo "hostprof rc=$?" >> $OUT/rc.txt ;;
    yago)   KGE_FORCE_SPLIT=1 timeout 300 python bench.py --workload rotate_yago310 --no-extras --no-cpu-baseline --no-parity --eval-queries 1024 > $OUT/bench_yago_split.json 2> $OUT/bench_yago_split.err; echo "yago rc=$?" >> $OUT/rc.txt ;;
    ref)    timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?" >> $OUT/rc.txt ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity --eval-queries 512 > $OUT/ncu_launch.log 2>&1
      echo "ncu launches rc=$?" >> $OUT/rc.txt
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"row_kernel_split|entity_kernel" -c 4 -o $OUT/prof_train \
        python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-parity --eval-queries 512 > $OUT/ncu_full.log 2>&1
      echo "ncu full rc=$?" >> $OUT/rc.txt ;;
  esac
done
cat $OUT/rc.txt
