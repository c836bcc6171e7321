#!/bin/bash
# Is the step kernel of the current default library the SAME machine code as at git revision <rev>?  Compiles csrc/env_step.cu of that revision
# (with the CURRENT header: appended struct fields must not change the code either) and compares every kernel's SASS instruction by instruction
# with the current build.  Usage: tools/check_sass_identical.sh [rev]     (default bc31fe5 = the source the round-1 GPU measurements were made with)
set -e
REV=${1:-bc31fe5}
cd "$(dirname "$0")/.."
T=$(mktemp -d)
for f in env_step.cu env_step_core.cuh common.cuh; do git show $REV:go2_rl_gym_b200/csrc/$f > $T/$f; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
nvcc $FLAGS -I go2_rl_gym_b200/csrc -c $T/env_step.cu -o $T/old.o 2>/dev/null
nvcc $FLAGS -c go2_rl_gym_b200/csrc/env_step.cu -o $T/new.o 2>/dev/null
for n in old new; do cuobjdump -sass $T/$n.o | grep -E "^\s+/\*[0-9a-f]{4,}\*/|Function :" | sed 's#/\* 0x[0-9a-f]* \*/##' > $T/$n.sass; done
python3 - $T $REV <<'PY'
import sys
def split(path):
    out, cur, order = {}, None, []
    for line in open(path):
        if "Function :" in line:
            cur = line.split("Function :")[1].strip(); out[cur] = []; order.append(cur)
        elif cur is not None: out[cur].append(line)
    return out, order
a, _ = split(sys.argv[1] + "/old.sass"); b, ob = split(sys.argv[1] + "/new.sass")
print(f"# step-kernel SASS of the working tree vs revision {sys.argv[2]} (sm_100a, -O3 -lineinfo)")
bad = 0
for k in ob:
    st = "identical" if a.get(k) == b[k] else ("new kernel" if k not in a else "DIFFERENT")
    bad += st == "DIFFERENT"
    print(f"{len(b[k]):6d} instructions  {st:10s}  {k[:110]}")
sys.exit(1 if bad else 0)
PY
