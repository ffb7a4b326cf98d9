run() { echo "=== $*"; env "$@" PCB_BENCH_DUMP_STEPS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 >/dev/null | grep "region" | awk '{printf "%s %s %s %s ", $1,$2,$3,$4; m=0; for(i=7;i<=NF;i++) if($i>m) m=$i; print "max_step=" m}'; }
run A=1
run PCB_MID_STEP_CHAIN=0
run PCB_NO_PDL=1
run PCB_NO_WGRAD_SIDE_STREAM=1
