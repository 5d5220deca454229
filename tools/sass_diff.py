import subprocess, hashlib, re, sys
def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d = {}; name = None; h = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = re.sub(r"Lb0E", "", m.group(1)); d[name] = hashlib.md5(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?)\s*/\* 0x", ln)
        if m and name:
            d[name].update(m.group(1).encode())
    return {k: v.hexdigest() for k, v in d.items()}
a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
changed = [k for k in a if k in b and a[k] != b[k]]
print("functions: before %d, after %d; identical %d; changed %d; removed %d; added %d" % (len(a), len(b), len(same), len(changed), len([k for k in a if k not in b]), len([k for k in b if k not in a])))
for k in changed: print("CHANGED", k[:100])
for k in a:
    if k not in b: print("REMOVED", k[:100])
for k in b:
    if k not in a: print("ADDED", k[:100])
