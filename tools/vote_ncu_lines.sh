ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:vote_kernel_grouped -s 2 -c 1 -o gpurun_out/vote_lines -f python tools/profile_vote.py 10000 50000 64 2 > gpurun_out/vote_lines.log 2>&1
tail -2 gpurun_out/vote_lines.log
