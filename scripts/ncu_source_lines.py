"""Per-source-line stall / instruction profile from an .ncu-rep whose CUDA view ncu cannot correlate (the GPU box builds under
another path): the SASS page of the report is joined, instruction by instruction, with `nvdisasm -g` of the SAME in-tree build.
usage: python scripts/ncu_source_lines.py <report.ncu-rep> <kernel-substring> [top_n]
(the library must be the build the report was taken from: the script checks that the opcode sequences agree)"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multi-source-lms-for-audio_b200", "libvqb_b200.so")


def disassemble(kernel):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
        for f in sorted(os.listdir(tmp)):
            out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
            heads = [i for i, l in enumerate(out.splitlines()) if l.startswith("//--------------------- .text.") and kernel in l]
            if heads:
                lines = out.splitlines()
                start = heads[0]
                end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- ")), len(lines))
                cur, dis = None, []
                for l in lines[start:end]:
                    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
                    if m:
                        cur = (os.path.basename(m.group(1)), int(m.group(2)))
                        continue
                    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
                    if m:
                        dis.append((m.group(2).strip(), cur))
                return dis
    raise SystemExit(f"kernel {kernel!r} not found in {LIB}")


def opcode(text):
    return re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    sass = [(r[1].strip(), int(r[ii] or 0), int(r[si] or 0)) for r in rows[2:] if len(r) > ii]
    dis = disassemble(kernel)
    if len(dis) != len(sass) or any(opcode(a[0]) != opcode(b[0]) for a, b in zip(dis, sass)):
        raise SystemExit(f"the report ({len(sass)} instructions) was not taken from this build ({len(dis)} instructions)")
    samples, insts = collections.Counter(), collections.Counter()
    for (_, src), (_, n, s) in zip(dis, sass):
        samples[src] += s
        insts[src] += n
    ts, ti = sum(samples.values()), sum(insts.values())
    cache = {}
    if "--waits" in sys.argv:
        # inlined helpers (vqb_ptx.cuh: the mbarrier wait loop) hide which ROLE waits: attribute their samples to the next
        # instruction that comes from another file, i.e. to the code that was waiting
        ctx_s, ctx_i = collections.Counter(), collections.Counter()
        srcs = [d[1] for d in dis]
        for i, ((_, src), (_, n, smp)) in enumerate(zip(dis, sass)):
            if src and src[0] == "vqb_ptx.cuh":
                j = i
                while j < len(srcs) and (srcs[j] is None or srcs[j][0] == "vqb_ptx.cuh"):
                    j += 1
                key = srcs[j] if j < len(srcs) else None
                ctx_s[key] += smp
                ctx_i[key] += n
        print("-- vqb_ptx.cuh samples by the code that follows the wait:")
        for key, v in ctx_s.most_common(14):
            text = ""
            if key:
                path = os.path.join(ROOT, "multi-source-lms-for-audio_b200", "csrc", key[0])
                if os.path.exists(path):
                    cache.setdefault(path, open(path).read().splitlines())
                    text = cache[path][key[1] - 1].strip()[:110]
            print(f"   before {key[0] if key else '?'}:{key[1] if key else 0:<6d} samples {100 * v / ts:5.1f} %  instructions {100 * ctx_i[key] / ti:5.1f} %  {text}")
    print(f"{rep}: {ts} samples, {ti} warp instructions")
    for (src, s) in samples.most_common(top):
        text = ""
        if src:
            path = os.path.join(ROOT, "multi-source-lms-for-audio_b200", "csrc", src[0])
            if os.path.exists(path):
                cache.setdefault(path, open(path).read().splitlines())
                text = cache[path][src[1] - 1].strip()[:100]
        where = f"{src[0]}:{src[1]}" if src else "?"
        print(f"{where:28s} samples {100 * s / ts:5.1f} %  instructions {100 * insts[src] / ti:5.1f} %  {text}")


if __name__ == "__main__":
    main()
