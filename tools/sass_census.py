"""SASS opcode census of libvlgba.so (cuobjdump -sass): which Blackwell-specific instructions the build actually contains,
and the FP64 / memory instruction mix of the hot kernels.  Usage: python tools/sass_census.py > profiles/sass_census_rNN.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bundleadjustmentmatlab_b200", "libvlgba.so")
INTEREST = ["DMMA", "UBLKCP", "UTMALDG", "UTMASTG", "UTCMMA", "SYNCS", "LDGSTS", "MUFU", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS",
            "SHFL", "BAR", "CALL", "LDL", "STL", "ATOM", "RED", "CCTL", "MEMBAR", "ERRBAR", "FENCE"]
HOT = ["k_stage1_cam<6, false>", "k_stage1_pt_tiled", "k_vinv_damp", "k_schur_chunk", "k_schur_fold<6>", "k_schur_blocks_heavy<6>",
       "k_schur_blocks_light<6>", "k_cluster_inverse<6>", "k_symv_lower", "k_pcg_persistent<6>", "k_chol_coop", "k_backsub_tiled<6>",
       "k_new_cost<6>", "k_sweep_pt_tiled<6>", "k_sweep_cam_ring<6, 2>"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    funcs = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    dem = {k: re.sub(r"^(void )?vlgba::", "", n).split("(")[0].replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true")
           for k, n in zip(funcs, names)}
    tot = collections.Counter()
    for c in funcs.values():
        tot.update(c)
    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(funcs)} functions, arch {', '.join(arch)}")
    print("whole library:", ", ".join(f"{k} {tot[k]}" for k in INTEREST if tot[k]))
    print("absent (as expected for an FP64 path: tcgen05 has no f64 kind, tiles move by 1-D bulk copies):",
          ", ".join(k for k in ("UTCMMA", "UTMALDG", "UTMASTG") if not tot[k]))
    print()
    for want in HOT:
        for k, c in funcs.items():
            if dem[k] == want or dem[k].startswith(want):
                n = sum(c.values())
                fp64 = c["DFMA"] + c["DMUL"] + c["DADD"] + c["DMMA"]
                print(f"{dem[k]}: {n} instructions, FP64 {fp64} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} DMMA {c['DMMA']}), MUFU {c['MUFU']}, "
                      f"LDG {c['LDG']} STG {c['STG']} LDS {c['LDS']} STS {c['STS']} LDGSTS {c['LDGSTS']} UBLKCP {c['UBLKCP']} SYNCS {c['SYNCS']} "
                      f"SHFL {c['SHFL']} BAR {c['BAR']} CALL {c['CALL']} LDL {c['LDL']} STL {c['STL']}")
                break


if __name__ == "__main__":
    main()
