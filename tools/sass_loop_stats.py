"""Static SASS statistics of a kernel (cuobjdump -sass): instruction count per pipe class between two
addresses, to steer instruction-count work without a GPU.
    python tools/sass_loop_stats.py k_dp3 [start_hex end_hex]
"""
import re, subprocess, sys, collections
ALU = ("LOP3", "SEL", "ISETP", "IADD3", "SHF", "VIMNMX", "VIADD", "VIADDMNMX", "PLOP3", "PRMT", "LEA", "IABS", "FMNMX", "MOV", "P2R", "R2P", "CS2R")
FMA = ("IMAD", "HFMA2", "FFMA", "FMUL", "FADD")
XU = ("BREV", "FLO", "POPC", "I2F", "F2I", "MUFU")
def cls(op):
    b = op.split(".")[0]
    if b in FMA: return "fma"
    if b in XU: return "xu"
    if b in ALU: return "alu"
    if b in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LD", "ST", "ATOMG", "RED", "LDC", "LDCU", "ATOMS"): return "lsu"
    if b in ("SHFL", "VOTE", "VOTEU", "CREDUX", "REDUX", "MATCH"): return "warp"
    if b in ("BRA", "BSSY", "BSYNC", "BREAK", "EXIT", "WARPSYNC", "NOP", "CALL", "RET", "BAR"): return "ctl"
    return "uni" if b.startswith("U") else "other"
def main():
    name = sys.argv[1]
    out = subprocess.run(["cuobjdump", "-sass", "falcon_b200/libfalcon_b200.so"], capture_output=True, text=True).stdout
    on = False; rows = []
    for l in out.splitlines():
        if "Function :" in l:
            on = name in l
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if on and m:
            rows.append((int(m.group(1), 16), m.group(3)))
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
    c = collections.Counter(cls(op) for a, op in rows if lo <= a <= hi)
    print(name, "instructions", sum(c.values()), dict(c))
main()
