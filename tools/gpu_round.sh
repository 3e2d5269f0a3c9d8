#!/bin/bash
# One gpurun session: GPU parity tests, bench of the schedules, launch list and one full ncu capture.
# usage (from the repo root on the GPU box): bash tools/gpu_round.sh <tag> [steps...]
set -u
TAG=${1:-r1}
shift || true
STEPS=${*:-"smoke tests bench history reference launches full"}
export ARTISB200_BENCH_CACHE=/tmp/bench_cache
mkdir -p gpurun_out
for step in $STEPS; do
  case $step in
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log ;;
    reference)
      timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?" ;;
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" ;;
    bench)
      timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?" ;;
    workloads)
      # the other BASELINE configurations (bench.py --workload): one bench line each; asym3d a second time with the
      # per-cell tables forced into windows (cell-batched tables)
      for w in ${WORKLOADS:-classic_1d3d gamma_3d50 asym3d}; do
        timeout ${WL_TIMEOUT:-900} python bench.py --workload $w --steps ${WL_STEPS:-2} --warmup ${WL_WARMUP:-1} > gpurun_out/${TAG}_bench_${w}.json 2> gpurun_out/${TAG}_bench_${w}.err; echo "workload $w rc=$?"
      done
      if [[ " ${WORKLOADS:-classic_1d3d gamma_3d50 asym3d} " == *" asym3d "* ]]; then
        ARTISB200_OPTS="table_budget_mb=${WINDOW_BUDGET_MB:-8192}" timeout ${WL_TIMEOUT:-900} python bench.py --workload asym3d --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_asym3d_windows.json 2> gpurun_out/${TAG}_bench_asym3d_windows.err; echo "workload asym3d (windows) rc=$?"
      fi ;;
    final)
      # the round's record: bench line, reference arm, DRAM traffic of every launch of one step, launch list, full captures
      timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
      timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
      timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/${TAG}_dram.csv python bench.py --one-step > /dev/null 2> gpurun_out/${TAG}_dram.err; echo "dram rc=$?"
      python tools/ncu_dram_traffic.py gpurun_out/${TAG}_dram.csv gpurun_out/${TAG}_dram_traffic.json | tee gpurun_out/${TAG}_dram_summary.txt
      ARTISB200_BENCH_NPACKETS=2000000 timeout 600 ncu --set full --clock-control none --import-source on \
        -k regex:k_wf_stage --launch-skip 18 --launch-count 6 -o gpurun_out/${TAG}_full -f \
        python bench.py --one-step > /dev/null 2> gpurun_out/${TAG}_full.err; echo "full rc=$?"
      ARTISB200_BENCH_NPACKETS=2000000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 600 \
        --log-file gpurun_out/${TAG}_launches.csv python bench.py --one-step > /dev/null 2> gpurun_out/${TAG}_launches.err; echo "launches rc=$?" ;;
    spectra)
      # the spectra / light-curve binning kernel on 1e7 synthetic final packets, the reference's own binning
      # timed beside it, and one full ncu capture of the kernel
      timeout 300 python tools/bench_spectra.py --packets 10000000 --out gpurun_out/${TAG}_spectra_bench.json > gpurun_out/${TAG}_spectra_bench.log 2>&1; echo "spectra rc=$?"
      timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bin_escaped -c 1 -f -o gpurun_out/${TAG}_spectra_full \
        python tools/bench_spectra.py --packets 4000000 --reference-replicas 0 --passes 1 --out gpurun_out/${TAG}_spectra_under_ncu.json > gpurun_out/${TAG}_spectra_ncu.log 2>&1; echo "spectra ncu rc=$?" ;;
    history)
      ARTISB200_OPTS="schedule=0" timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_history.json 2> gpurun_out/${TAG}_bench_history.err; echo "history rc=$?" ;;
    variants)
      for v in "wf_tail=0" "wf_tail=65536" "wf_rsteps_thin=2" "wf_rsteps_thick=1" "wf_rsteps_thick=32"; do
        ARTISB200_OPTS="$v" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_${v}.json 2> gpurun_out/${TAG}_bench_${v}.err; echo "$v rc=$?"
      done ;;
    tune)
      # TUNE="<lib suffix>:<options>;..." e.g. TUNE=":wf_masteps=1;_b6:wf_masteps=2"
      IFS=';' read -ra CASES <<< "${TUNE:-:}"
      for cs in "${CASES[@]}"; do
        suf="${cs%%:*}"; opts="${cs#*:}"
        name="${TAG}_tune${suf}_$(echo "$opts" | tr ',=' '__')"
        ARTISB200_LIB_SUFFIX="$suf" ARTISB200_OPTS="$opts" timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${name}.json 2> gpurun_out/${name}.err; echo "tune [$suf] [$opts] rc=$?"
      done ;;
    dram)
      # DRAM bytes of every launch of ONE full-size step (three metrics = one replay pass per launch)
      timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/${TAG}_dram.csv python bench.py --one-step > /dev/null 2> gpurun_out/${TAG}_dram.err; echo "dram rc=$?"
      python tools/ncu_dram_traffic.py gpurun_out/${TAG}_dram.csv gpurun_out/${TAG}_dram_traffic.json | tee gpurun_out/${TAG}_dram_summary.txt ;;
    thinfull)
      # full ncu capture of two launches of the detailed r-packet stage kernel of a 2e6-packet step
      ARTISB200_BENCH_NPACKETS=2000000 timeout 1500 ncu --set full --clock-control none --import-source on \
        --kernel-name-base demangled -k 'regex:k_wf_stage<\(int\)1>' --launch-skip 2 --launch-count 2 -o gpurun_out/${TAG}_thinfull -f \
        python bench.py --one-step > /dev/null 2> gpurun_out/${TAG}_thinfull.err; echo "thinfull rc=$?" ;;
    launches)
      ARTISB200_BENCH_NPACKETS=2000000 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 600 \
        --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.json 2> gpurun_out/${TAG}_launches.err; echo "launches rc=$?" ;;
    full)
      # one early (large) launch of each stage kernel of a 1e6-packet step
      ARTISB200_BENCH_NPACKETS=${FULL_NPACKETS:-2000000} ARTISB200_OPTS="${FULL_OPTS:-}" timeout 1500 ncu --set full --clock-control none --import-source on \
        -k regex:k_wf_stage --launch-skip ${FULL_SKIP:-8} --launch-count ${FULL_COUNT:-4} -o gpurun_out/${TAG}_full -f \
        python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2> gpurun_out/${TAG}_full.err; echo "full rc=$?" ;;
  esac
done
ls -la gpurun_out | tail -20
