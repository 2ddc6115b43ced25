"""Offline SASS loop census: for one kernel of libsdfmesh.so, list the backward branches (loops) and the opcode mix of
each loop body.  Usage: python tools/sass_loops.py k_project [min_len]"""
import collections
import re
import subprocess
import sys

so = "bevy-signed-distance-mesh-generation_b200/libsdfmesh.so"
kern = sys.argv[1]
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 100
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = next(b for b in blocks if re.match(r"_ZN3sdm\d+" + kern + r"E", b))
ins = []
for line in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
print(kern, "instructions:", len(ins))
ALU = ("FMNMX", "FSETP", "FSEL", "IADD", "ISETP", "LOP3", "SEL", "IMAD.MOV", "MOV", "PLOP3", "SHF", "LEA", "IABS", "FCHK", "POPC", "FLO", "BREV", "I2F", "F2I", "PRMT", "VIADD", "IMNMX", "VIMNMX", "FMNMX3")
def cls(op):
    o = op.split()[0]
    if o.startswith("@"):
        o = op.split()[1]
    if o.startswith(("FADD", "FMUL", "FFMA")): return "fma"
    if o.startswith("IMAD"): return "imad"
    if o.startswith("MUFU"): return "xu"
    if o.startswith(("LD", "ST", "ATOM", "RED")): return "mem"
    if o.startswith(("SHFL", "VOTE", "REDUX", "MATCH")): return "warp"
    if o.startswith(("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BAR", "NOP", "BREAK", "JMP")): return "ctl"
    if o.startswith(ALU): return "alu"
    return "other:" + o
for i, (a, op) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)", op)
    if not m: continue
    t = int(m.group(1), 16)
    if t < a and t in addr_idx:
        j = addr_idx[t]
        n = i - j + 1
        if n < min_len: continue
        c = collections.Counter(cls(o) for _, o in ins[j:i + 1])
        ops = collections.Counter((o.split()[1] if o.startswith("@") else o.split()[0]) for _, o in ins[j:i + 1])
        print(f"loop {t:#x}..{a:#x} len {n}: " + ", ".join(f"{k}={v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))
        print("     top ops: " + ", ".join(f"{k}={v}" for k, v in ops.most_common(14)))
