# light ncu pass: duration, issue utilisation and instruction-fetch stalls of the neighbor kernel (scratch tool)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for spec in "$@"; do
  set -- $spec; label=$1; shift
  env A=1 "$@" ncu --metrics $M --clock-control none -k regex:neighbors_kernel -c 1 -s 3 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -E "^\"[0-9]" | awk -F'","' -v l=$label '{printf "%s %s %s\n", l, $(NF-2), $NF}' | sed 's/"//g'
done
