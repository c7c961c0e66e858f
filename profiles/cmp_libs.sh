# usage: bash profiles/cmp_libs.sh lib1.so lib2.so ...   (variants built with drl-tetris_b200/build.py --out; parity first, then the bench legs)
for lib in "$@"; do
export B2T_LIB=$PWD/drl-tetris_b200/b200tetris/$lib
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "world_model or rollout_matches_oracle or c2_4096" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --leg placements 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'step %.4f ms  %.3e/s frac %.3f | placements %.4f ms %.3e/s'%(d['ms_per_step'], d['value'], d['roofline']['frac'], d['placements']['ms_per_expand'], d['placements']['value']))"
done
