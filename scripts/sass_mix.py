"""Development aid: pipe mix of a SASS instruction range (fmaheavy: IDP / IMAD*, alu: the rest of the integer ops),
usage: python scripts/sass_mix.py <object> <function regex> [first last]   (without a range: every loop with IDPs)"""
import re, sys
sys.path.insert(0, __import__('os').path.dirname(__file__))
from sass_loops import functions  # noqa: E402


def mix(ls):
    fma = sum(1 for l in ls if re.match(r'(@!?U?P\d\s+)?(IDP|IMAD|FFMA|FMUL|VIADD\.16)', l.strip()))
    lsu = sum(1 for l in ls if re.match(r'(@!?U?P\d\s+)?(LDS|STS|LDG|STG|LD\.|ST\.|ATOM|RED)', l.strip()))
    uni = sum(1 for l in ls if re.match(r'(@!?U?P\d\s+)?(UMOV|ULDC|LDCU|UIADD|USHF|ULOP|UISETP|USEL|UPRMT|R2UR|LDC)', l.strip()))
    ctl = sum(1 for l in ls if re.match(r'(@!?U?P\d\s+)?(BRA|BSSY|BSYNC|VOTE|WARPSYNC|NOP|BAR|EXIT|ELECT|CALL|RET)', l.strip()))
    return dict(n=len(ls), fmaheavy=fma, lsu=lsu, uniform=uni, control=ctl, alu=len(ls) - fma - lsu - uni - ctl)


if __name__ == '__main__':
    obj, pat = sys.argv[1], sys.argv[2]
    for name, ls in functions(obj):
        if not re.search(pat, name):
            continue
        if len(sys.argv) > 4:
            print(mix(ls[int(sys.argv[3]):int(sys.argv[4]) + 1]))
        else:
            # fast paths: from a loop head to the first VOTE after >= 30 IDPs
            for i, l in enumerate(ls):
                if 'BRA' in l:
                    m = re.search(r'0x([0-9a-f]+)', l)
                    if m:
                        t = int(m.group(1), 16) // 16
                        if t < i and 100 < i - t < 900 and sum('IDP' in b for b in ls[t:i]) >= 30:
                            idp = 0
                            for k in range(t, i):
                                idp += 'IDP' in ls[k]
                                if 'VOTE' in ls[k] and idp >= 30:
                                    print(f'loop {t}->{i}: head+fast path {t}..{k}', mix(ls[t:k + 1]))
                                    break
                            else:
                                print(f'loop {t}->{i}: whole', mix(ls[t:i + 1]))
        break
