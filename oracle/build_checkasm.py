#!/usr/bin/env python3
"""Builds the REFERENCE'S OWN parity harness (tests/checkasm) against libdav1d_cuda.so.

TEST INFRASTRUCTURE ONLY.  Nothing from the reference is committed: this script copies
tests/checkasm/{checkasm.c,checkasm.h,mc.c,itx.c,ipred.c} and the three DSP templates plus cpu.c
from the reference tree (default /root/reference) into the git-ignored oracle/_ref/checkasm/,
applies the three small patches a maintainer would make to hook a new backend in
(SURVEY.md section 7, step 1):

  (i)   a new entry in checkasm's cpus[] table (tests/checkasm/checkasm.c:96-123) for the flag bit
        DAV1D_CUDA_CPU_FLAG = 1 << 30 (x86 uses bits 0-5, src/x86/cpu.h:31-41);
  (ii)  dav1d_init_cpu() (src/cpu.c:54-69) reports that bit, so that checkasm's
        check_cpu_flag() (checkasm.c:530-553) runs a pass for it;
  (iii) the tail of dav1d_{mc,itx,intra_pred}_dsp_init_{8,16}bpc (src/mc_tmpl.c:948-956,
        src/itx_tmpl.c:270-283, src/ipred_tmpl.c:767-773) calls dav1d_cuda_*_dsp_init_* when the
        bit is set - the same pattern as the per-arch *_dsp_init_x86(c);

trims checkasm's test list to the three components this backend overrides, and links the result
with -ldav1d_cuda.  checkasm then compares every function of the CUDA tables against the C
templates exactly as it does for an assembly version.  Usage:

  oracle/_ref/checkasm/checkasm --test=mc_8bpc 1        (seed 1; also itx_*, ipred_*, *_16bpc)
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(HERE, "_ref", "checkasm")
LIBDIR = os.path.join(ROOT, "dav1d-mirror_b200")


def patch(path, subs):
    s = open(path).read()
    for pat, rep, count in subs:
        s2, n = re.subn(pat, rep, s, flags=re.S)
        if n != count:
            raise SystemExit(f"{path}: pattern {pat!r} matched {n} times, expected {count}")
        s = s2
    open(path, "w").write(s)


def main():
    if not os.path.isdir(os.path.join(REF, "tests", "checkasm")):
        raise SystemExit(f"{REF}: no reference tree")
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(os.path.join(OUT, "src"))
    os.makedirs(os.path.join(OUT, "tests", "checkasm"))
    for f in ("checkasm.c", "checkasm.h", "mc.c", "itx.c", "ipred.c"):
        shutil.copy(os.path.join(REF, "tests", "checkasm", f), os.path.join(OUT, "tests", "checkasm", f))
    for f in ("mc_tmpl.c", "itx_tmpl.c", "ipred_tmpl.c", "cpu.c"):
        shutil.copy(os.path.join(REF, "src", f), os.path.join(OUT, "src", f))

    flag = "(1u << 30)"
    ck = os.path.join(OUT, "tests", "checkasm", "checkasm.c")
    patch(ck, [
        # (i) the new "cpu" flag
        (r'(\{ "AVX-512 \(Ice Lake\)", "avx512icl", DAV1D_X86_CPU_FLAG_AVX512ICL \},\n)',
         r'\1    { "CUDA (sm_100a)",     "cuda",      %s },\n' % flag, 1),
        # only the components this backend overrides
        (r'    \{ "msac", checkasm_check_msac \},\n    \{ "pal", checkasm_check_pal \},\n'
         r'    \{ "refmvs", checkasm_check_refmvs \},\n', '', 1),
        (r'    \{ "(cdef|filmgrain|loopfilter|looprestoration)_(8|16)bpc", checkasm_check_\w+ \},\n', '', 8),
        # no assembly: no AVX warm-up / cpuid helpers to call
        (r'#if ARCH_X86_64\n(\s+)void checkasm_warmup_avx2\(void\);', r'#if ARCH_X86_64 && HAVE_ASM\n\1void checkasm_warmup_avx2(void);', 1),
        (r'#if ARCH_X86\n(\s+)unsigned checkasm_init_x86\(char \*name\);', r'#if ARCH_X86 && HAVE_ASM\n\1unsigned checkasm_init_x86(char *name);', 1),
    ])
    # (ii) the flag is "detected"
    patch(os.path.join(OUT, "src", "cpu.c"), [
        (r'(COLD void dav1d_init_cpu\(void\) \{\n)', r'\1    dav1d_cpu_flags |= %s; /* CUDA backend */\n' % flag, 1),
    ])
    # (iii) the init tails
    hook = ('\n    {{ /* CUDA backend (libdav1d_cuda.so), same pattern as *_dsp_init_x86(c) */\n'
            '        extern void bitfn(dav1d_cuda_{name}_dsp_init)({proto});\n'
            '        if (dav1d_get_cpu_flags() & %s) bitfn(dav1d_cuda_{name}_dsp_init)({args});\n'
            '    }}\n}}\n' % flag)
    for fname, name, proto, args in (("mc_tmpl.c", "mc", "void *", "c"),
                                     ("itx_tmpl.c", "itx", "void *, int", "c, bpc"),
                                     ("ipred_tmpl.c", "intra_pred", "void *", "c")):
        p = os.path.join(OUT, "src", fname)
        s = open(p).read()
        i = s.rstrip().rfind("}")
        s = '#include "src/cpu.h"\n' + s[:i] + hook.format(name=name, proto=proto, args=args).lstrip("\n").replace("    {", "\n    {", 1)
        open(p, "w").write(s)

    cflags = ["-std=c99", "-O2", "-march=x86-64-v3", "-fwrapv", "-D_GNU_SOURCE", "-w",
              "-I" + os.path.join(HERE, "ref_cfg"), "-I" + OUT, "-I" + REF, "-I" + os.path.join(REF, "include"),
              "-I" + os.path.join(REF, "include", "dav1d"), "-I" + os.path.join(REF, "tests")]
    objs = []

    def cc(src, out, extra=()):
        o = os.path.join(OUT, out)
        subprocess.check_call(["gcc", *cflags, *extra, "-c", src, "-o", o])
        objs.append(o)

    for bpc in (8, 16):
        for t in ("mc_tmpl", "itx_tmpl", "ipred_tmpl"):
            cc(os.path.join(OUT, "src", t + ".c"), f"{t}_{bpc}.o", [f"-DBITDEPTH={bpc}"])
        for t in ("mc", "itx", "ipred"):
            cc(os.path.join(OUT, "tests", "checkasm", t + ".c"), f"check_{t}_{bpc}.o", [f"-DBITDEPTH={bpc}"])
        cc(os.path.join(REF, "src", "ipred_prepare_tmpl.c"), f"ipred_prepare_{bpc}.o", [f"-DBITDEPTH={bpc}"])
    for t in ("itx_1d", "tables", "wedge", "scan"):
        cc(os.path.join(REF, "src", t + ".c"), t + ".o")
    cc(os.path.join(OUT, "src", "cpu.c"), "cpu.o")
    cc(ck, "checkasm.o")
    # the only library symbol cpu.c wants besides libc (src/log.c needs the whole decoder context)
    stub = os.path.join(OUT, "stubs.c")
    open(stub, "w").write("void dav1d_log(void *c, const char *fmt, ...) { (void)c; (void)fmt; }\n")
    cc(stub, "stubs.o")
    exe = os.path.join(OUT, "checkasm")
    subprocess.check_call(["gcc", "-o", exe, *objs, "-L" + LIBDIR, "-ldav1d_cuda", "-lm", "-lpthread",
                           "-Wl,-rpath,$ORIGIN/../../../dav1d-mirror_b200"])
    print("built", exe)


if __name__ == "__main__":
    main()
