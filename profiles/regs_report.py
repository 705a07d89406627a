"""Register / spill report of the kernels in a ptxas -v log (csrc/ptxas_*.log)."""
import re
import subprocess
import sys

txt = open(sys.argv[1]).read()
blocks = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                    r"(\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
for name, st, ss, sl, regs in blocks:
    dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace('void amgb::', '').split('(')[0]
    if len(sys.argv) > 2 and sys.argv[2] not in dem:
        continue
    print("%-60s regs %3s stack %4s spill st/ld %4s/%4s" % (dem, regs, st, ss, sl))
