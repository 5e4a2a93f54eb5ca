#!/bin/bash
# Static profile of the built kernels (no GPU needed): ptxas resource usage and SASS instruction counts per function.
# usage: bash tools/static_profile.sh > profiles/r2_static.txt      (after python xrslam_b200/build.py)
cd "$(dirname "$0")/../xrslam_b200/csrc" || exit 1
echo "# ptxas -v (registers, spills, stack, shared memory) -- nvcc $(nvcc --version | grep release | sed 's/.*release //')"
for f in backend_solve.cu backend_marg.cu backend_preint.cu frontend.cu; do
  flags="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v"
  [ $f = frontend.cu ] && flags="$flags -fmad=false"
  nvcc $flags -c $f -o /tmp/static_$$.o 2>&1 | grep -E "Function properties|spill|Used" | sed -e 's/ptxas info    : //' -e 's/_INTERNAL_[0-9a-f_]*//' | cut -c1-220
done
rm -f /tmp/static_$$.o
echo
echo "# SASS instruction counts per function (cuobjdump -sass of xrslam_b200/build/*.o)"
cd ../build || exit 1
for o in backend_solve.o backend_marg.o backend_preint.o frontend.o; do
  echo "== $o"
  cuobjdump -sass $o 2>/dev/null | awk '/Function :/ {fn=$3; n[fn]=0} {n[fn]++} /DMMA/ {d[fn]++} /UBLKCP/ {u[fn]++} /UTMALDG/ {t[fn]++} /BAR\.SYNC/ {b[fn]++} /SYNCS/ {m[fn]++} /DFMA/ {f[fn]++} /DADD|DMUL/ {a[fn]++} /LDL/ {l[fn]++} /STL/ {s[fn]++} /LDG/ {g[fn]++} /LDS/ {h[fn]++}
    END {for (k in n) if (k != "") printf "%-90s lines %6d  DFMA %5d  DADD/DMUL %5d  DMMA %4d  UBLKCP %2d  UTMALDG %2d  mbarrier %2d  BAR.SYNC %3d  LDG %4d  LDS %4d  LDL %4d  STL %4d\n", substr(k,1,90), n[k], f[k], a[k], d[k], u[k], t[k], m[k], b[k], g[k], h[k], l[k], s[k]}' | sort
done
