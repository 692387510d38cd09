#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
run() {
  echo "== $*"
  env "$@" timeout 300 python tools/prof_forward.py --tokens 1048576 --reps 1 2>&1 | tail -2
  env "$@" timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:rq_forward -s 1 -c 1 --csv \
    --log-file $OUT/l2sweep_x.csv python tools/prof_forward.py --tokens 1048576 --reps 1 > /dev/null 2>&1
  python tools/traffic_json.py $OUT/l2sweep_x.csv 1048576; grep hit_rate $OUT/l2sweep_x.csv | cut -d, -f14-
}
timeout 300 python -m pytest tests -m gpu -x -q -k "lockstep or 2b_kat" 2>&1 | tail -3
run RQAE_LOCKSTEP=0
run RQAE_LOCKSTEP=1
run RQAE_LOCKSTEP=1 RQAE_L2_HOT=0.6
run RQAE_LOCKSTEP=1 RQAE_L2_HOT=0.4
run RQAE_LOCKSTEP=1 RQAE_L2_HOT=0
