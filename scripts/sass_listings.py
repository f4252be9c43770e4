"""Per-kernel SASS listings of the sm_100a cubins (`cuobjdump -sass` on the in-tree .so files, encodings stripped) and a
mnemonic table.  Usage: python scripts/sass_listings.py   (writes profiles/sass/*.sass and profiles/sass/README.md)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "colossalai_b200", "kernel", "_build")
OUT = os.path.join(ROOT, "profiles", "sass")
# (output name, library, substrings that must all appear in the mangled function name)
KERNELS = [
    ("gemm_tcgen05_2cta_bk128", "libcb200_gemm.so", ["gemm_tcgen05_2cta_kernel", "Li128ELi3ELb0E"]),
    ("gemm_fp8_tcgen05_2cta_bk128", "libcb200_gemm.so", ["gemm_tcgen05_2cta_kernel", "Li128ELi3ELb1E"]),
    ("grouped_gemm_nt", "libcb200_grouped_gemm.so", ["grouped_gemm_2cta_kernel", "Li0E"]),
    ("grouped_gemm_tn_wgrad", "libcb200_grouped_gemm.so", ["grouped_gemm_2cta_kernel", "Li2E"]),
    ("flash_fwd_d128", "libcb200_attn.so", ["flash_fwd_kernel", "Li128ELb0E"]),
    ("flash_fwd_d128_window_alibi", "libcb200_attn.so", ["flash_fwd_kernel", "Li128ELb1E"]),
    ("flash_bwd_d128", "libcb200_attn.so", ["flash_bwd_kernel", "Li128E"]),
    ("fused_ag_gemm_2cta", "libcb200_comm.so", ["fused_gemm_2cta_kernel", "Li0E"]),
    ("fused_gemm_rs_ar_2cta", "libcb200_comm.so", ["fused_gemm_2cta_kernel", "Li1E"]),
    ("ulysses_a2a", "libcb200_comm.so", ["ulysses_a2a_kernel"]),
    ("paged_decode_mma_bf16_d128", "libcb200_infer.so", ["paged_decode_mma_kernel", "bfloat16", "Li128E"]),
    ("moe_ep_push_rows_bf16", "libcb200_moe.so", ["ep_push_rows_kernel", "bfloat16"]),
]
COLS = [("UTC*MMA", r"\bUTC\w*MMA"), ("UTMALDG", r"\bUTMALDG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("SYNCS", r"\bSYNCS"), (".SYS", r"\.SYS\b"), ("LDGMC", r"\bLDGMC"), ("REDG", r"\bREDG"), ("HMMA", r"\bHMMA"),
        ("LDSM", r"\bLDSM"), ("LDGSTS", r"\bLDGSTS")]


def functions(lib):
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, lib)], capture_output=True, text=True).stdout
    out, name, buf = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = buf
            name, buf = m.group(1), []
        elif name is not None:
            buf.append(line)
    if name:
        out[name] = buf
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    cache, rows = {}, []
    for out_name, lib, keys in KERNELS:
        if lib not in cache:
            cache[lib] = functions(lib)
        hits = [n for n in cache[lib] if all(k in n for k in keys)]
        if not hits:
            print(f"[sass] no function matches {keys} in {lib}", file=sys.stderr)
            continue
        body = []
        for line in cache[lib][hits[0]]:
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\*", line)
            if m:
                body.append(f"/*{m.group(1)}*/  {m.group(2)} ;")
        with open(os.path.join(OUT, out_name + ".sass"), "w") as f:
            f.write(f"// {hits[0]}\n" + "\n".join(body) + "\n")
        text = "\n".join(body)
        rows.append([out_name, str(len(body))] + [str(len(re.findall(rx, text))) for _, rx in COLS])
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# SASS listings of the tcgen05 / TMA / tensor-core / peer-memory kernels\n\n"
                "`python scripts/sass_listings.py`: `cuobjdump -sass` of the sm_100a cubins in `colossalai_b200/kernel/_build/*.so` "
                "(one file per kernel, encoding words stripped).  Mnemonic legend (B200_PROFILING.md): `UTC*MMA` = tcgen05.mma, "
                "`UTMALDG` = TMA tile load, `LDTM` / `STTM` = tcgen05.ld / tcgen05.st, `SYNCS` = mbarrier ops, `.SYS`-scoped LDG/STG = "
                "peer-memory / cross-GPU flag traffic, `LDGMC` = multimem.ld_reduce (in-switch reduction), `REDG` = red.global.add "
                "(dQ / remote dK,dV accumulation), `HMMA` = mma.sync, `LDSM` = ldmatrix, `LDGSTS` = cp.async.\n\n")
        f.write("| kernel | lines | " + " | ".join(c for c, _ in COLS) + " |\n|---|---|" + "---|" * len(COLS) + "\n")
        for r in rows:
            f.write("| " + " | ".join(r) + " |\n")
    print(f"[sass] wrote {len(rows)} listings to {OUT}")


if __name__ == "__main__":
    main()
