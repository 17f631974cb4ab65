"""Development helper (GPU box): per-opcode issue rates (lane-level chain steps per second)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynaalign_b200._lib import lib  # noqa: E402

NAMES = ["IADD(add folded)", "VIADDMNMX", "VIMNMX3", "PRMT", "SEL", "ISETP+SEL", "LOP3(folded)", "SHF", "IMAD", "HSET2+LOP3", "HADD2/HFMA2",
         "VIADDMNMX.U16x2", "VIMNMX.S16x2+pred(+VIADD)", "VIMNMX3.S16x2", "ISETP+@P VIADD", "SHFL.UP", "LDS(+3 int)", "POPC+IADD",
         "vibmax_s32 (ISETP+2SEL+IADD)", "HSET2 only", "HSET2 + ISUB", "VIMNMX3 + PRMT", "IMAD + PRMT", "VIADDMNMX + ISETP + @P VIADD",
         "IADD3 3-input", "HSET2 + IMAD", "PRMT + SEL", "VIMNMX3 + IMAD", "ISETP + @P MOV",
         "IMAD.HI", "VIADDMNMX.S16x2 + IADD", "VIADDMNMX.S16x2 + 2 IADD", "VIADDMNMX.S16x2 + IMAD.HI", "VIADDMNMX.S16x2 + IMAD", "IADD + IMAD",
         "LDS.128 + 4 VIADDMNMX.S16x2"]
import sys as _s
if len(_s.argv) > 1:
    NAMES = [(n if i >= int(_s.argv[1]) else None) for i, n in enumerate(NAMES)]
L = C.CDLL(lib()._name)
L.dyna_probe_op.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_void_p]
for op, name in enumerate(NAMES):
    if name is None:
        continue
    r = C.c_double(0)
    rc = L.dyna_probe_op(op, C.byref(r), None)
    print("op %2d %-32s %7.2f T chain-steps/s  (%.1f lanes/clk/SM @1.96GHz x148)" % (op, name, r.value / 1e12, r.value / 148 / 1.96e9), flush=True)
