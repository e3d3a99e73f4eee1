"""Summarise k_llms_b200/csrc/build.log (ptxas -v): registers, stack, spills per kernel."""
import re
import subprocess
import sys

log = open(sys.argv[1] if len(sys.argv) > 1 else "k_llms_b200/csrc/build.log").read()
blocks = re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]
for b in blocks:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void kc::", "")
    st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    rg = re.search(r"Used (\d+) registers", b)
    print(f"{dem:45s} regs={rg.group(1):>3s} stack={st.group(1):>4s} spill={st.group(2)}/{st.group(3)}")
