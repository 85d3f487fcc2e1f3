"""Dev tool: SASS evidence for profiles/ -- per kernel the opcode histogram and the TMA / mbarrier / cluster / async-store
instructions (cuobjdump -sass of the objects __graft_entry__.build() leaves under pyloo_b200/lib/obj).

    python tests/dev/sass_excerpts.py > profiles/r2_sass_excerpts.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OBJ = os.path.join(ROOT, "pyloo_b200", "lib", "obj")
WANT = [("b2l_tile.o", "loo_tile_kernelILi8ELb0"), ("b2l_tile.o", "loo_tile_kernelILi8ELb1"), ("b2l_tile.o", "tile_merge_kernel"),
        ("b2l_split_tail.o", "psis_tail_kernelILi8ELi1ELi8"), ("b2l_split_stream.o", "psis_stream_kernelILi256ELi16ELi0")]
SPECIAL = re.compile(r"UTMALDG|UTMASTG|UBLKCP|SYNCS|STAS|UCGABAR|ATOMG|REDG|MEMBAR|FENCE|CCTL|ERRBAR|BAR\.")

def main():
    for obj, key in WANT:
        path = os.path.join(OBJ, obj)
        txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        blocks = re.split(r"\n\s*Function : ", txt)
        for b in blocks[1:]:
            name = b.split("\n", 1)[0].strip()
            if key not in name:
                continue
            ins = re.findall(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)([^;]*);", b)
            hist = collections.Counter(op.split(".")[0] for _, op, _ in ins)
            print(f"==== pyloo_b200/lib/obj/{obj} :: {name}  ({len(ins)} SASS instructions)")
            print("opcode histogram (top 24): " + ", ".join(f"{k} {v}" for k, v in hist.most_common(24)))
            print("TMA / mbarrier / cluster / async-store / atomic instructions:")
            seen = collections.Counter()
            for addr, op, rest in ins:
                if SPECIAL.search(op):
                    seen[op.split(".")[0]] += 1
                    if seen[op.split(".")[0]] <= 12:
                        print(f"  /*{addr}*/  {op}{rest}")
            print("  (counts: " + ", ".join(f"{k} {v}" for k, v in seen.items()) + ")")
            print()

if __name__ == "__main__":
    main()
