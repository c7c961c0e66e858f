# usage: bash profiles/dev/steplibs.sh lib1.so lib2.so ... : the step leg of bench.py for each library variant (3 runs each)
for lib in "$@"; do
export B2T_LIB=$PWD/drl-tetris_b200/b200tetris/$lib
for r in 1 2 3; do
python bench.py --steps 30 --warmup 5 --leg step 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', 'step %.4f ms  %.3e/s frac %.3f'%(d['ms_per_step'], d['value'], d['roofline']['frac']))"
done
done
