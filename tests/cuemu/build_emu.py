"""TEST INFRASTRUCTURE ONLY.  Builds libxrvio_emu.so: the product's unmodified .cu sources compiled by g++ against the
stand-in cuda_runtime.h of this directory and linked with the SIMT emulator (cuemu.cpp), so the kernels can be executed
on a machine without a GPU.  Two textual rewrites are applied to copies of the sources, nothing else:
    kernel<<<grid, block, smem, stream>>>(args);   ->  cuemu::launch(grid, block, smem, [&] { kernel(args); });
    extern __shared__ T name[];                    ->  T *name = (T *)cuemu::dyn_smem();
    __device__ T name[...];  (namespace scope)     ->  static T name[...];
The product library (xrslam_b200/lib/libxrvio.so) is never built or replaced by this."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'xrslam_b200', 'csrc')
SAN = os.environ.get('XRVIO_EMU_SANITIZE') == '1'      # AddressSanitizer build: out-of-bounds accesses of the kernels and the host glue
TSAN = os.environ.get('XRVIO_EMU_SANITIZE') == 'thread'   # ThreadSanitizer build: data races of the multi-threaded host glue (use with CUEMU_REPLAY: no fibers)
OUT = os.path.join(HERE, '_build_asan' if SAN else ('_build_tsan' if TSAN else '_build'))
LIB = os.path.join(OUT, 'libxrvio_emu.so')
SOURCES = ['frontend.cu', 'backend_preint.cu', 'backend_solve.cu', 'backend_marg.cu', 'capi_ops.cu', 'capi_session.cu', 'xrslam_shim.cu']


def _split_top(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip()); cur = ''
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite(src):
    # namespace-scope device variables are per translation unit under nvcc (no relocatable device code)
    src = re.sub(r'(?m)^(__device__|__constant__)\s+((?:unsigned |const )*[A-Za-z_]\w*(?:\s+[A-Za-z_]\w*)*\s+\w+\s*(?:\[[^\]]*\])?\s*(?:=[^;]*)?;)', r'static \2', src)
    src = re.sub(r'extern\s+__shared__\s+([A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;', r'\1 *\2 = (\1 *)cuemu::dyn_smem();', src)
    out, pos = '', 0
    while True:
        i = src.find('<<<', pos)
        if i < 0:
            out += src[pos:]
            break
        j = i
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] in '_:<>, '):       # kernel name, possibly templated
            j -= 1
        name = src[j:i].strip()
        k = src.index('>>>', i)
        cfg = _split_top(src[i + 3:k])
        while len(cfg) < 3:
            cfg.append('0')
        p = src.index('(', k)
        depth, q = 0, p
        while True:
            if src[q] == '(':
                depth += 1
            elif src[q] == ')':
                depth -= 1
                if depth == 0:
                    break
            q += 1
        args = src[p + 1:q]
        out += src[pos:j] + ' cuemu::launch(%s, %s, %s, [&]() { %s(%s); })' % (cfg[0], cfg[1], cfg[2], name, args)
        pos = q + 1
    return out


def build(verbose=False):
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, 'cuemu.cpp'), os.path.join(HERE, 'prof.cpp'), os.path.join(HERE, 'include', 'cuda_runtime.h'),
                                                                 os.path.join(ROOT, 'include', 'xrvio.h'), __file__]
    if os.path.exists(LIB) and os.path.getmtime(LIB) > max(os.path.getmtime(s) for s in srcs):
        return LIB
    gen = os.path.join(OUT, 'gen', 'csrc')
    os.makedirs(gen, exist_ok=True)
    for f in os.listdir(CSRC):                     # headers and .inl files may hold launches too
        with open(os.path.join(CSRC, f)) as fh:
            text = fh.read()
        with open(os.path.join(gen, f if not f.endswith('.cu') else f[:-3] + '.cpp'), 'w') as fh:
            fh.write(rewrite(text))
    os.makedirs(os.path.join(OUT, 'include'), exist_ok=True)
    # the sources include "../../include/xrvio.h" relative to csrc/: mirror that layout under _build/
    for h in ('xrvio.h', 'xrslam_compat.h'):
        with open(os.path.join(ROOT, 'include', h)) as fh, open(os.path.join(OUT, 'include', h), 'w') as oh:
            oh.write(fh.read())
    objs, procs = [], []
    flags = ['-O2', '-g' if os.environ.get('XRVIO_EMU_DEBUGINFO') == '1' else '-g0', '-std=c++17', '-fPIC', '-w', '-ffp-contract=off', '-I', os.path.join(HERE, 'include')]
    if SAN:
        flags = ['-O1', '-g', '-fsanitize=address', '-fno-omit-frame-pointer'] + flags[2:]
    if TSAN:
        flags = ['-O1', '-g', '-fsanitize=thread', '-fno-omit-frame-pointer'] + flags[2:]
    for s in SOURCES:
        cpp = os.path.join(gen, s[:-3] + '.cpp')
        obj = os.path.join(OUT, s[:-3] + '.o')
        objs.append(obj)
        cmd = ['g++'] + flags + ['-c', cpp, '-o', obj]
        if verbose:
            print(' '.join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    pobj = os.path.join(OUT, 'prof.o')
    procs.append(('prof.cpp', subprocess.Popen(['g++'] + flags + ['-c', os.path.join(HERE, 'prof.cpp'), '-o', pobj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs.append(pobj)
    obj = os.path.join(OUT, 'cuemu.o')
    procs.append(('cuemu.cpp', subprocess.Popen(['g++'] + flags + ['-c', os.path.join(HERE, 'cuemu.cpp'), '-o', obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs.append(obj)
    bad = False
    for s, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            bad = True
            sys.stderr.write('g++ failed for %s:\n%s\n' % (s, o.decode()[:6000]))
    if bad:
        raise RuntimeError('emulated library build failed')
    subprocess.check_call(['g++', '-shared', '-o', LIB] + objs + ['-lpthread', '-ldl'] + (['-fsanitize=address'] if SAN else []) + (['-fsanitize=thread'] if TSAN else []))
    return LIB


if __name__ == '__main__':
    print(build(verbose=True))
