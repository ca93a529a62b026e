"""Dynamic instruction mix of one kernel from two `ncu --set full --import-source on` reports (SASS page).

    python tools/instruction_mix.py <kernel-regex> before.ncu-rep after.ncu-rep > profiles/<tag>_instruction_mix.md

Groups the executed warp instructions by opcode class (the grouping of profiles/r01_instruction_mix.md)."""
import collections
import csv
import subprocess
import sys

GROUPS = [
    ("FP64 arithmetic (DFMA, DMUL, DADD, DSETP)", ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")),
    ("register moves (MOV, IMAD.MOV, UMOV, CS2R)", ("MOV", "IMAD.MOV", "UMOV", "CS2R")),
    ("integer / address (IMAD, IADD3, LEA, LOP3, SHF, VIADD, ISETP, SEL, PLOP3, IMNMX)",
     ("IMAD", "IADD3", "IADD", "LEA", "LOP3", "SHF", "VIADD", "ISETP", "SEL", "PLOP3", "IMNMX", "VIMNMX", "IABS",
      "UIADD3", "ULOP3", "UISETP", "USEL", "USHF", "UIMAD", "ULEA", "UPLOP3", "POPC", "FLO", "BREV", "PRMT", "R2UR",
      "UPRMT", "R2P", "P2R", "LOP", "ULOP", "UFLO", "UPOPC", "VIADDMNMX")),
    ("FP32-pipe helpers (FSEL, FSETP, FFMA, FMUL, FADD, HFMA2, F2F, I2F, F2I, FCHK)",
     ("FSEL", "FSETP", "FFMA", "FMUL", "FADD", "HFMA2", "F2F", "I2F", "F2I", "FCHK", "FMNMX", "HADD2", "I2FP", "F2FP",
      "FRND", "HMUL2")),
    ("control flow (BRA, BSSY, BSYNC, BREAK, CALL, RET, WARPSYNC, VOTE, SHFL, NOP, EXIT)",
     ("BRA", "BSSY", "BSYNC", "BREAK", "CALL", "RET", "WARPSYNC", "VOTE", "SHFL", "NOP", "EXIT", "BRX", "BAR", "VOTEU",
      "BMOV", "JMP", "NANOSLEEP", "YIELD", "ELECT", "MATCH", "REDUX", "ENDCOLLECTIVE", "BPT", "ACQBULK", "ERRBAR",
      "DEPBAR", "MEMBAR", "CCTL", "WARPSYNC.ALL")),
    ("memory (LDG, STG, LDL, STL, LDS, STS, LDC, LDCU, ATOM)",
     ("LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "LDCU", "ATOM", "ATOMS", "ATOMG", "RED", "LD", "ST", "S2R",
      "S2UR", "CS2UR", "LDSM", "STSM")),
    ("MUFU", ("MUFU",)),
]


def mix(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    iS, iI, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    agg = collections.Counter()
    tot = thr = 0
    for r in rows[hi + 1:]:
        if r and r[0] in ("Address", "Kernel Name"):   # a second listing of the kernel: keep the first only
            break
        if len(r) <= iT or not r[iI].isdigit():
            continue
        toks = r[iS].split()
        op = next((t for t in toks if not t.startswith("@") and not t.startswith("{")), "")
        n = int(r[iI])
        tot += n
        thr += int(r[iT])
        base = op.split(".")[0]
        key = "IMAD.MOV" if op.startswith("IMAD.MOV") else base
        g = next((name for name, ops in GROUPS if key in ops), "other")
        agg[g] += n
    return agg, tot, thr


def main():
    kernel, before, after = sys.argv[1:4]
    a, ta, tha = mix(before, kernel)
    b, tb, thb = mix(after, kernel)
    print("| group | before (M warp-inst) | share | after (M warp-inst) | share |\n|---|---|---|---|---|")
    for name in [g[0] for g in GROUPS] + ["other"]:
        print("| %s | %.0f | %.1f%% | %.0f | %.1f%% |" % (name, a[name] / 1e6, 100 * a[name] / ta, b[name] / 1e6, 100 * b[name] / tb))
    print("| **total** | **%.0f** | | **%.0f** (%.0f%%) | |" % (ta / 1e6, tb / 1e6, 100 * tb / ta))
    print("| active lanes per instruction | %.1f | | %.1f | |" % (tha / ta, thb / tb))


if __name__ == "__main__":
    main()
