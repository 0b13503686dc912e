# A/B runs of the grid sweep on one GPU (tools, not part of the product): name, then environment settings
run() { name=$1; shift; env "$@" PCM_STAGE_TIMELINE=gpurun_out/${TAG:-ab}_tl_$name.json python bench.py --workload sweep > gpurun_out/${TAG:-ab}_sweep_$name.json 2> gpurun_out/${TAG:-ab}_sweep_$name.err; echo "$name exit $?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG:-ab}_sweep_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), 'seq/s', round(d['mean_iou'],6))"; }
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name $(echo $envs | tr ',' ' ')
done
