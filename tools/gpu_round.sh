#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, the bench line, A/B over the tuning knobs of the two fuse
# kernels, the ncu launch list of a short bench run and ncu --set full captures of the two fuse kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a [ab] [notest] [noncu]'
# Everything lands in gpurun_out/<tag>_*; nothing printed under ncu is a bench value.
TAG=${1:-run}
shift
OUT=gpurun_out
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1

if [[ " $* " != *" notest "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi

timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json

if [[ " $* " == *" ab "* ]]; then
  # triples MSL_SCAN_STAGES _ MSL_APPLY_CTAS _ MSL_APPLY_ILP (override the list with AB_CFGS="0_2_4 0_4_2 ...")
  for cfg in ${AB_CFGS:-0_2_4 0_4_2 0_3_2 0_4_1}; do
    IFS=_ read S A I <<< "$cfg"
    MSL_SCAN_STAGES=$S MSL_APPLY_CTAS=$A MSL_APPLY_ILP=$I timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
      > $OUT/${TAG}_ab_$cfg.json 2>> $OUT/${TAG}_bench.err
    python tools/ab_line.py $OUT/${TAG}_ab_$cfg.json "scan_stages=$S apply_ctas=$A apply_ilp=$I"
  done
fi

if [[ " $* " != *" noncu "* ]]; then
  # launch list (per-launch times are cold-cache and serialised: shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
  python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1
  head -30 $OUT/${TAG}_launches_summary.txt
  # full captures of the two fuse kernels (one launch each, a frame in the steady state of the stream)
  for kn in ${NCU_KERNELS:-k_fuse_one}; do
    SKIP=40
    [[ $kn == k_sp_* ]] && SKIP=2
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s $SKIP -c 1 -f -o $OUT/${TAG}_$kn \
      python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline >> $OUT/${TAG}_ncu_bench.log 2>&1
    python tools/ncu_brief.py $OUT/${TAG}_$kn.ncu-rep > $OUT/${TAG}_${kn}_brief.txt 2>&1
    cat $OUT/${TAG}_${kn}_brief.txt
  done
fi
