#!/usr/bin/env python3
"""Generate octopuszk_b200/csrc/chains.cuh: the carry-chain blocks of the 256-bit field arithmetic.

Each block is described ONCE as a list of PTX instructions and emitted twice: as one non-volatile inline-asm
statement for nvcc, and as the same instruction list run through the carry-flag emulation of ptx_arith.cuh for the
CPU-side tests (tests/host_arith_check.cc).  Generating both from one list keeps the CPU-tested algorithm and the
PTX text from drifting apart."""
import os
import re

out = []

COP = {'add.cc.u32': 'add_cc', 'addc.cc.u32': 'addc_cc', 'addc.u32': 'addc', 'sub.cc.u32': 'sub_cc',
       'subc.cc.u32': 'subc_cc', 'subc.u32': 'subc', 'mad.lo.cc.u32': 'mad_lo_cc', 'madc.lo.cc.u32': 'madc_lo_cc',
       'madc.hi.cc.u32': 'madc_hi_cc', 'madc.hi.u32': 'madc_hi', 'mul.lo.u32': 'mul_lo'}


def block(name, params, instrs, doc):
    """params: (cname, kind, count), kind in {'io', 'in', 'out'}; instrs: (op, dst, srcs) with operand tokens such
    as 'e3', 'bi', '0' (literal zero) or 'm' (block-local temporary)."""
    order = []
    for cname, kind, count in params:
        if kind in ('io', 'out'):
            order += [(cname, i, kind, count) for i in range(count)]
    for cname, kind, count in params:
        if kind == 'in':
            order += [(cname, i, kind, count) for i in range(count)]
    idx = {(c, i): n for n, (c, i, _, _) in enumerate(order)}
    counts = {c: n for c, _, n in params}

    def split(t):
        m = re.match(r'([a-z]+)(\d*)$', t)
        return m.group(1), int(m.group(2)) if m.group(2) else 0

    def tok(t):
        if t in ('0', 'm'):
            return t
        return '%%%d' % idx[split(t)]

    def ctok(t):
        if t == '0':
            return '0u'
        if t == 'm':
            return 'm'
        c, i = split(t)
        return c if counts[c] == 1 else '%s[%d]' % (c, i)

    sig = []
    for cname, kind, count in params:
        if count == 1:
            sig.append(('uint32_t& ' if kind != 'in' else 'uint32_t ') + cname)
        else:
            sig.append(('uint32_t* ' if kind != 'in' else 'const uint32_t* ') + cname)
    uses_m = any('m' in (ins[1],) + tuple(ins[2]) for ins in instrs)
    out.append('// ' + doc)
    out.append('OZK_D void %s(%s) {' % (name, ', '.join(sig)))
    out.append('#if defined(__CUDACC__)')
    lines = ['{ .reg .u32 m;'] if uses_m else []
    lines += ['%s %s, %s;' % (op, tok(dst), ', '.join(tok(s) for s in srcs)) for op, dst, srcs in instrs]
    if uses_m:
        lines.append('}')
    out.append('    asm(')
    out.extend(['        "%s\\n\\t"' % l for l in lines])
    outs, ins_ = [], []
    for cname, i, kind, count in order:
        ref = cname if count == 1 else '%s[%d]' % (cname, i)
        if kind == 'io':
            outs.append('"+r"(%s)' % ref)
        elif kind == 'out':
            outs.append('"=r"(%s)' % ref)
        else:
            ins_.append('"r"(%s)' % ref)
    out.append('        : ' + ', '.join(outs))
    out.append('        : ' + ', '.join(ins_) + ');')
    out.append('#else')
    out.append('    using namespace ptx;')
    if uses_m:
        out.append('    uint32_t m;')
    out.extend(['    %s = %s(%s);' % (ctok(dst), COP[op], ', '.join(ctok(s) for s in srcs)) for op, dst, srcs in instrs])
    out.append('#endif')
    out.append('}')
    out.append('')


# r = a + b
ins = [('add.cc.u32', 'r0', ('a0', 'b0'))] + [('addc.cc.u32', 'r%d' % i, ('a%d' % i, 'b%d' % i)) for i in range(1, 7)]
ins += [('addc.u32', 'r7', ('a7', 'b7'))]
block('add8', [('r', 'out', 8), ('a', 'in', 8), ('b', 'in', 8)], ins,
      'r = a + b over 8 limbs; the carry out of limb 7 is dropped (callers guarantee there is none)')

# r = a - b with borrow mask
ins = [('sub.cc.u32', 'r0', ('a0', 'b0'))] + [('subc.cc.u32', 'r%d' % i, ('a%d' % i, 'b%d' % i)) for i in range(1, 8)]
ins += [('subc.u32', 'bw', ('0', '0'))]
block('sub8', [('r', 'out', 8), ('bw', 'out', 1), ('a', 'in', 8), ('b', 'in', 8)], ins,
      'r = a - b over 8 limbs; bw = 0xffffffff when a < b, else 0')

# Montgomery round, part A: T += a * b_i (rounds 1..7)
ins = [('add.cc.u32', 'e0', ('e0', 'o1'))]
for j in (0, 2, 4):
    ins.append(('madc.lo.cc.u32', 'o%d' % j, ('a%d' % (j + 1), 'bi', 'o%d' % (j + 2))))
    ins.append(('madc.hi.cc.u32', 'o%d' % (j + 1), ('a%d' % (j + 1), 'bi', 'o%d' % (j + 3))))
ins.append(('madc.lo.cc.u32', 'o6', ('a7', 'bi', '0')))
ins.append(('madc.hi.u32', 'o7', ('a7', 'bi', '0')))
ins.append(('mad.lo.cc.u32', 'e0', ('a0', 'bi', 'e0')))
ins.append(('madc.hi.cc.u32', 'e1', ('a0', 'bi', 'e1')))
for j in (2, 4, 6):
    ins.append(('madc.lo.cc.u32', 'e%d' % j, ('a%d' % j, 'bi', 'e%d' % j)))
    ins.append(('madc.hi.cc.u32', 'e%d' % (j + 1), ('a%d' % j, 'bi', 'e%d' % (j + 1))))
ins.append(('addc.u32', 'o7', ('o7', '0')))
block('mont_round_ab', [('e', 'io', 8), ('o', 'io', 8), ('a', 'in', 8), ('bi', 'in', 1)], ins,
      "T += a * bi for rounds 1..7.  On entry o is last round's limb-0 accumulator (o[0] == 0, o[1] folds into e[0]) "
      "and e last\n// round's limb-1 accumulator; on exit e is the limb-0 and o the limb-1 accumulator of this round.")

# Montgomery round, part B: m = T[0] * np; T += m * p
ins = [('mul.lo.u32', 'm', ('e0', 'np'))]
ins.append(('mad.lo.cc.u32', 'o0', ('p1', 'm', 'o0')))
ins.append(('madc.hi.cc.u32', 'o1', ('p1', 'm', 'o1')))
for j in (2, 4, 6):
    ins.append(('madc.lo.cc.u32', 'o%d' % j, ('p%d' % (j + 1), 'm', 'o%d' % j)))
    ins.append(('madc.hi.cc.u32', 'o%d' % (j + 1), ('p%d' % (j + 1), 'm', 'o%d' % (j + 1))))
ins.append(('mad.lo.cc.u32', 'e0', ('p0', 'm', 'e0')))
ins.append(('madc.hi.cc.u32', 'e1', ('p0', 'm', 'e1')))
for j in (2, 4, 6):
    ins.append(('madc.lo.cc.u32', 'e%d' % j, ('p%d' % j, 'm', 'e%d' % j)))
    ins.append(('madc.hi.cc.u32', 'e%d' % (j + 1), ('p%d' % j, 'm', 'e%d' % (j + 1))))
ins.append(('addc.u32', 'o7', ('o7', '0')))
block('mont_round_mp', [('e', 'io', 8), ('o', 'io', 8), ('p', 'in', 8), ('np', 'in', 1)], ins,
      'm = e[0] * np mod 2^32; T += m * p.  Afterwards e[0] == 0 (T is divisible by 2^32).')

# Montgomery reduction of a 16-limb value, shift step between rounds: drop the (zero) low limb of the limb-0 accumulator `o`,
# fold its limb 1 into e[0], move the rest down by two limbs and bring the next input limb t in at the top.
ins = [('add.cc.u32', 'e0', ('e0', 'o1'))]
for j in range(6):
    ins.append(('addc.cc.u32', 'o%d' % j, ('o%d' % (j + 2), '0')))
ins.append(('addc.cc.u32', 'o6', ('t', '0')))
ins.append(('addc.u32', 'o7', ('0', '0')))
block('redc_shift', [('e', 'io', 8), ('o', 'io', 8), ('t', 'in', 1)], ins,
      "Shift between two rounds of a stand-alone Montgomery reduction: e[0] += o[1]; o = (o >> 64) + carry, with the next input\n"
      "// limb t entering at o[6].  Afterwards e is the limb-0 and o the limb-1 accumulator.")

# 8-limb add / sub with carry (borrow) in and out, for 16-limb values handled as two halves
ins = [('add.cc.u32', 'x', ('cin', 'ones'))]          # CC = cin
ins += [('addc.cc.u32', 'r%d' % i, ('a%d' % i, 'b%d' % i)) for i in range(8)]
ins += [('addc.u32', 'cout', ('0', '0'))]
block('add8c', [('r', 'out', 8), ('cout', 'out', 1), ('x', 'out', 1), ('a', 'in', 8), ('b', 'in', 8), ('cin', 'in', 1), ('ones', 'in', 1)], ins,
      'r = a + b + cin over 8 limbs, cout = carry out (cin, cout in {0,1}; ones must be 0xffffffff; x is scratch)')
ins = [('sub.cc.u32', 'x', ('0', 'bin'))]             # borrow flag = bin
ins += [('subc.cc.u32', 'r%d' % i, ('a%d' % i, 'b%d' % i)) for i in range(8)]
ins += [('subc.u32', 'bout', ('0', '0'))]
block('sub8b', [('r', 'out', 8), ('bout', 'out', 1), ('x', 'out', 1), ('a', 'in', 8), ('b', 'in', 8), ('bin', 'in', 1)], ins,
      'r = a - b - bin over 8 limbs, bout = 0xffffffff when a borrow leaves limb 7 else 0 (bin in {0,1}; x is scratch)')

HDR = '''// GENERATED by tools/gen_chains.py -- do not edit.
// Carry-chain blocks of the 256-bit field arithmetic.  Each block is ONE non-volatile inline-asm statement (so the
// carry flag never crosses a statement the compiler could reorder, and NVVM sees a handful of pure asm calls per
// multiplication instead of hundreds of volatile ones); the #else branch is the same instruction list run through
// the carry-flag emulation of ptx_arith.cuh, which is what the CPU-side tests execute.
#pragma once
#include "ptx_arith.cuh"
namespace ozk {
namespace chain {
'''
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'octopuszk_b200', 'csrc', 'chains.cuh')
open(path, 'w').write(HDR + '\n'.join(out) + '}  // namespace chain\n}  // namespace ozk\n')
print('wrote', path)
