# usage: bash profiles/enumcmp.sh lib1.so ... : duration / instructions / SIMT efficiency of enumerate_kernel (BOARDS=n to change the 65 536 default)
for lib in "$@"; do
export B2T_LIB=$PWD/drl-tetris_b200/b200tetris/$lib
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum -k regex:enumerate --csv --log-file /tmp/enum_$lib.csv python bench.py --steps 3 --warmup 3 --leg placements --boards ${BOARDS:-65536} --cuda-profile placements > /dev/null 2>&1
echo "$lib $(grep enumerate_kernel /tmp/enum_$lib.csv | awk -F'","' '{printf "%s=%s  ", $(NF-2), $NF}' | tr -d '"')"
done
