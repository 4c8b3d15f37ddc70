echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; echo "v1 quota: $(cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null) $(cat /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>/dev/null)"
nproc; python -c "import os; print(len(os.sched_getaffinity(0)))"
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo
cat /proc/loadavg
python tools/time_e2e.py 100 2>&1 | tail -7
cat /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo
cat /proc/loadavg
