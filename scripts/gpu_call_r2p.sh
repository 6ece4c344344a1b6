# 1 GPU: knob sweep on the headline workload (value only): lanes, WHILE-body unroll, K4 batch, blocks per SM and launch
mkdir -p gpurun_out
: > gpurun_out/r2p_sweep.log
run() { echo "== $*" >> gpurun_out/r2p_sweep.log; env "$@" python bench.py --steps 5 --warmup 3 --skip-cpu --skip-e2e ${LANES:+--lanes $LANES} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['config']['lanes'])" >> gpurun_out/r2p_sweep.log; }
LANES= run X=0
LANES=3 run X=0
LANES=5 run X=0
LANES=6 run X=0
LANES=8 run X=0
LANES= run STK_ECC_UNROLL=2
LANES= run STK_ECC_UNROLL=8
LANES= run STK_WARP_BATCH=2
LANES= run STK_WARP_BATCH=1
LANES= run STK_ECC_BLOCKS_PER_SM=2
LANES=6 run STK_ECC_BLOCKS_PER_SM=1
LANES= run STK_ECC_RIM_WEIGHT=8
LANES= run STK_ECC_RIM_WEIGHT=12
cat gpurun_out/r2p_sweep.log
