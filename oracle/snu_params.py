"""Write the npbparams.h that SNU_NPB/NPB3.3-OMP-C/sys/setparams.c would
generate for CG (class table: sys/setparams.c write_cg_info).  Test
infrastructure: only oracle/Makefile's `snu` target uses it."""
import sys

TABLE = {  # na, nonzer, niter, shift
    "S": (1400, 7, 15, "10.0"), "W": (7000, 8, 15, "12.0"),
    "A": (14000, 11, 15, "20.0"), "B": (75000, 13, 75, "60.0"),
    "C": (150000, 15, 75, "110.0"), "D": (1500000, 21, 100, "500.0"),
}
cls = sys.argv[1].upper()
na, nonzer, niter, shift = TABLE[cls]
print(f"""#define CLASS '{cls}'
#define NA {na}
#define NONZER {nonzer}
#define NITER {niter}
#define SHIFT {shift}
#define RCOND 1.0e-1
#define CONVERTDOUBLE false
#define COMPILETIME "oracle"
#define NPBVERSION "3.3.1"
#define CS1 "gcc"
#define CS2 "gcc"
#define CS3 "-lm"
#define CS4 "-I../common"
#define CS5 "-O3 -fopenmp"
#define CS6 "-O3 -fopenmp"
#define CS7 "randdp"
""")
