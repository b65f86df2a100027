"""Rewrites the three CUDA constructs g++ cannot parse so that the kernel sources compile against tests/emu/emu.h:

  kernel<T...><<<grid, block, smem, stream>>>(args)   ->  emu::launch_k(grid, block, smem, kernel<T...>, args)
  extern __shared__ [__align__(n)] T name[];          ->  T *name = (T *)emu::dyn_smem();          (ends at a guard page)
  __shared__ [__align__(n)] T name[N];                ->  T (&name)[N] = *emu::static_smem<T[N]>();    (ends at a guard page)
  asm("rcp.approx.ftz.f64 ..." / "rsqrt.approx.ftz.f64 ...")  ->  emu::rcp_approx_f64 / emu::rsqrt_approx_f64
  asm("{ setp.gt|lt.f64 p, x, 0; selp.f64 r, x, 0, p; }")      ->  r = x > 0 ? x : 0   (pos_part / neg_part of sph_math.cuh)
  asm("{ mov.b64 {lo,hi}, x; setp.ge|lt.s32 p, hi, 0; selp.f64 r, x, 0, p; }")  ->  the same by the sign bit (pos_part_s / neg_part_s)

Everything else (kernel bodies, launch logic, the C ABI) is compiled as written.  TEST INFRASTRUCTURE ONLY.
"""
import os
import re
import sys


def _match_forward(s, i, open_c, close_c):
    """s[i] == open_c; returns the index of the matching close_c."""
    depth = 0
    k = i
    while k < len(s):
        c = s[k]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return k
        k += 1
    raise ValueError("unbalanced %s%s" % (open_c, close_c))


def _split_top(s):
    out, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    out.append("".join(cur).strip())
    return out


def rewrite_launches(src):
    out = []
    pos = 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        # kernel expression, backwards: optional template argument list, then the (qualified) identifier
        k = i
        while k > 0 and src[k - 1].isspace():
            k -= 1
        if src[k - 1] == ">":
            depth = 0
            while k > 0:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        name = src[k:i].strip()
        j = src.index(">>>", i)
        cfg = _split_top(src[i + 3:j])
        while len(cfg) < 3:
            cfg.append("0")
        p = j + 3
        while src[p].isspace():
            p += 1
        assert src[p] == "(", "launch without argument list near: " + src[i - 40:i + 40]
        q = _match_forward(src, p, "(", ")")
        args = src[p + 1:q].strip()
        out.append(src[pos:k])
        out.append("emu::launch_k(%s, %s, %s, %s%s)" % (cfg[0], cfg[1], cfg[2], name, (", " + args) if args else ""))
        pos = q + 1
    return "".join(out)


_EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w\s]+?)\s+(\w+)\s*\[\s*\]\s*;")
_ASM_RCP = re.compile(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;"\s*:\s*"=d"\((\w+)\)\s*:\s*"d"\((\w+)\)\);')
_ASM_RSQ = re.compile(r'asm\("rsqrt\.approx\.ftz\.f64 %0, %1;"\s*:\s*"=d"\((\w+)\)\s*:\s*"d"\((\w+)\)\);')


_ASM_CLAMP = re.compile(r'asm\("\{ \.reg \.pred p; setp\.(gt|lt)\.f64 p, %1, 0d0+; selp\.f64 %0, %1, 0d0+, p; \}"\s*:\s*"=d"\((\w+)\)\s*:\s*"d"\((\w+)\)\);')


# the same selections by the sign bit (pos_part_s / neg_part_s): modelled on the bit pattern, as the instruction sees it
_ASM_CLAMP_S = re.compile(r'asm\("\{ \.reg \.pred p; \.reg \.b32 lo, hi; mov\.b64 \{lo, hi\}, %1; setp\.(ge|lt)\.s32 p, hi, 0; selp\.f64 %0, %1, 0d0+, p; \}"\s*:\s*"=d"\((\w+)\)\s*:\s*"d"\((\w+)\)\);')


_STATIC_SHARED = re.compile(r"(?<![\w])(?<!extern )__shared__\s+(?:__align__\((\d+)\)\s+)?([\w:\s]+?)\s+(\w+\s*(?:\[[^;]*\])?(?:\s*,\s*\w+\s*(?:\[[^;]*\])?)*)\s*;")


def rewrite_static_shared(src):
    """`__shared__ T a[N][M], b;` -> references to guarded storage (emu::static_smem): an index past the end of a
    statically sized shared array faults like one past the dynamic allocation does."""
    def repl(m):
        align, typ, decls = m.group(1), m.group(2).strip(), m.group(3)
        out = []
        for d in _split_top(decls):
            mm = re.match(r"(\w+)\s*(.*)$", d.strip(), re.S)
            name, dims = mm.group(1), mm.group(2).strip()
            al = ("__attribute__((aligned(%s))) " % align) if align else ""
            out.append("typedef %s emu_shared_t_%s %s%s; static emu_shared_t_%s &%s = *emu::static_smem<emu_shared_t_%s>();"
                       % (typ, name, dims, (" " + al) if al else "", name, name, name))
        return " ".join(out)
    return _STATIC_SHARED.sub(repl, src)


# packed FP32 pairs of sm_100 (sph_math.cuh): a 64-bit value holding two floats, lo first
_ASM_PACK = re.compile(r'asm\("mov\.b64 %0, \{%1, %2\};"\s*:\s*"=l"\((\w+)\)\s*:\s*"f"\((\w+)\),\s*"f"\((\w+)\)\);')
_ASM_UNPACK = re.compile(r'asm\("mov\.b64 \{%0, %1\}, %2;"\s*:\s*"=f"\((\w+)\),\s*"=f"\((\w+)\)\s*:\s*"l"\((\w+)\)\);')
_ASM_F32X2_2 = re.compile(r'asm\("(sub|mul)\.rn\.f32x2 %0, %1, %2;"\s*:\s*"=l"\((\w+)\)\s*:\s*"l"\((\w+)\),\s*"l"\((\w+)\)\);')
_ASM_F32X2_3 = re.compile(r'asm\("fma\.rn\.f32x2 %0, %1, %2, %3;"\s*:\s*"=l"\((\w+)\)\s*:\s*"l"\((\w+)\),\s*"l"\((\w+)\),\s*"l"\((\w+)\)\);')


_ASM_STS16 = re.compile(r'asm volatile\("st\.shared\.u16 \[%0\], %1;"\s*:\s*:\s*"r"\((\w+)\),\s*"h"\((\w+)\)\s*:\s*"memory"\);')


def transform(src):
    src = rewrite_launches(src)
    # shared-window addresses: offsets into the launch's dynamic shared memory
    src = src.replace("__cvta_generic_to_shared(", "emu::shared_offset(")
    src = _ASM_STS16.sub(lambda m: "*reinterpret_cast<unsigned short *>(emu::dyn_smem_ptr + %s) = %s;" % (m.group(1), m.group(2)), src)
    src = _ASM_PACK.sub(lambda m: "%s = emu::f32x2_pack(%s, %s);" % m.groups(), src)
    src = _ASM_UNPACK.sub(lambda m: "emu::f32x2_unpack(%s, %s, %s);" % (m.group(3), m.group(1), m.group(2)), src)
    src = _ASM_F32X2_2.sub(lambda m: "%s = emu::f32x2_%s(%s, %s);" % (m.group(2), m.group(1), m.group(3), m.group(4)), src)
    src = _ASM_F32X2_3.sub(lambda m: "%s = emu::f32x2_fma(%s, %s, %s);" % m.groups(), src)
    src = rewrite_static_shared(src)
    src = _EXTERN_SHARED.sub(lambda m: "%s *%s = (%s *)emu::dyn_smem();" % (m.group(1), m.group(2), m.group(1)), src)
    src = _ASM_RCP.sub(lambda m: "%s = emu::rcp_approx_f64(%s);" % (m.group(1), m.group(2)), src)
    src = _ASM_RSQ.sub(lambda m: "%s = emu::rsqrt_approx_f64(%s);" % (m.group(1), m.group(2)), src)
    src = _ASM_CLAMP.sub(lambda m: "%s = (%s %s 0.0) ? %s : 0.0;" % (m.group(2), m.group(3), ">" if m.group(1) == "gt" else "<", m.group(3)), src)
    src = _ASM_CLAMP_S.sub(lambda m: "%s = (((long long)__double_as_longlong(%s) %s 0) ? %s : 0.0);" %
                           (m.group(2), m.group(3), ">=" if m.group(1) == "ge" else "<", m.group(3)), src)
    if "asm(" in src or "asm volatile" in src:
        raise ValueError("inline PTX the emulator has no model for")
    return src


def main(csrc, gen):
    os.makedirs(gen, exist_ok=True)
    made = []
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cu", ".cuh")):
            continue
        text = open(os.path.join(csrc, f)).read()
        dst = os.path.join(gen, f[:-3] + ".cpp" if f.endswith(".cu") else f)
        new = '#include "emu.h"\n' + transform(text)
        if not os.path.exists(dst) or open(dst).read() != new:
            with open(dst, "w") as o:
                o.write(new)
        made.append(dst)
    return made


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
