"""Development aid: lengths and instruction mix of the loops of the PEE kernels (from cuobjdump -sass of the built
object), to compare instruction counts of the step loops before spending GPU time.
usage: python scripts/sass_loops.py [object] [name filter] ; python scripts/sass_loops.py dump <object> <function regex> <first> <last>"""
import re, subprocess, sys


def functions(obj):
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    out = []
    for f in re.split(r'\n\s+Function : ', txt)[1:]:
        name = f.split('\n', 1)[0]
        lines = [re.sub(r'\s*/\*.*$', '', re.sub(r'^\s+/\*[0-9a-f]+\*/\s+', '', l)) for l in f.split('\n')
                 if re.match(r'^\s+/\*[0-9a-f]{4,}\*/', l)]
        out.append((name, lines))
    return out


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "dump":
    obj, pattern, first, last = sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    for name, ls in functions(obj):
        if re.search(pattern, name):
            for i in range(first, min(last, len(ls) - 1) + 1):
                print(i, ls[i])
            break
    sys.exit(0)

if __name__ == '__main__':
    obj = sys.argv[1] if len(sys.argv) > 1 else '/root/repo/codec_tcc_b200/lib/peeb_pee2.o'
    flt = sys.argv[2] if len(sys.argv) > 2 else r'ItLi256ELi3E'
    for name, lines in functions(obj):
        if not re.search(flt, name):
            continue
        print(name[:60], 'total', len(lines))
        for i, l in enumerate(lines):
            if 'BRA' in l:
                m = re.search(r'0x([0-9a-f]+)', l)
                if m:
                    t = int(m.group(1), 16) // 16
                    if t < i and i - t > 60:
                        body = lines[t:i + 1]
                        if sum('IDP' in b for b in body) == 0:
                            continue
                        print(f"  loop {t}->{i} len {i-t+1} IDP {sum('IDP' in b for b in body)} MOV {sum('MOV' in b for b in body)} SEL {sum(b.startswith('SEL') for b in body)}")
