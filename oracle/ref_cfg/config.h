/* Hand-written build configuration for compiling the reference's C templates
 * (no asm, both bit-depth classes) straight from /root/reference.
 * TEST INFRASTRUCTURE ONLY - see oracle/README.md. */
#ifndef ORACLE_REF_CONFIG_H
#define ORACLE_REF_CONFIG_H
#define ARCH_X86 1
#define ARCH_X86_64 1
#define ARCH_X86_32 0
#define ARCH_AARCH64 0
#define ARCH_ARM 0
#define ARCH_LOONGARCH 0
#define ARCH_LOONGARCH64 0
#define ARCH_PPC64LE 0
#define ARCH_RISCV 0
#define ARCH_RV32 0
#define ARCH_RV64 0
#define HAVE_ASM 0
#define HAVE_AVX512ICL 0
#define CONFIG_8BPC 1
#define CONFIG_16BPC 1
#define CONFIG_LOG 1
#define TRIM_DSP_FUNCTIONS 0
#define ENDIANNESS_BIG 0
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_C11_GENERIC 1
#define HAVE_UNISTD_H 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_DLSYM 1
#define HAVE_PTHREAD_GETAFFINITY_NP 1
#define HAVE_PTHREAD_SETAFFINITY_NP 1
#define STACK_ALIGNMENT 16
#endif
