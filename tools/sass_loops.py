"""Developer tool: list the loops of a kernel in a cuobjdump -sass dump with their instruction mix.
usage: python tools/sass_loops.py /tmp/all.sass <kernel-name-substring> [--dump]"""
import re, sys
from collections import Counter
txt = open(sys.argv[1]).read()
pat = sys.argv[2]
dump = "--dump" in sys.argv
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, len(ins), "instructions")
    for addr, op in ins:
        m = re.search(r'BRA\s+(?:.*?)0x([0-9a-f]+)', op)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [(a, o) for a, o in ins if tgt <= a <= addr]
            ops = [re.sub(r'^@!?U?P\d+\s+', '', o).split()[0].split('.')[0] for a, o in body]
            c = Counter(ops)
            fp64 = c['DFMA'] + c['DMUL'] + c['DADD'] + c['DSETP']
            print("  loop %#x..%#x: %d instr, FP64 %d (DFMA %d DMUL %d DADD %d DSETP %d), MUFU %d, LDS %d, other %d" % (
                tgt, addr, len(body), fp64, c['DFMA'], c['DMUL'], c['DADD'], c['DSETP'], c['MUFU'], c['LDS'], len(body) - fp64))
            if dump:
                for a, o in body:
                    print("     %#x %s" % (a, o))
