"""Compare the SASS of every kernel of two builds of libptb200.so, function by function (addresses and
encodings stripped, names demangled): which kernels changed between a GPU-validated commit and now.

    python performance-test_b200/tools/sass_compare.py old/libptb200.so new/libptb200.so
"""
import re, subprocess, sys, collections
def dump(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []; continue
        if cur is None: continue
        if line.startswith('Fatbin') or line.startswith('====') or re.match(r'^(arch|code version|host|compile_size|identifier) =', line) or 'code for sm_' in line or '.target' in line: cur = None if line.startswith('Fatbin') else cur; continue
        if re.match(r"^\s*/\* 0x[0-9a-f]+ \*/\s*$", line): continue   # second encoding line
        line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line)               # encoding
        line = re.sub(r"^\s*/\*[0-9a-f]{4}\*/", "", line)             # address
        funcs[cur].append(line.rstrip())
    return funcs
def demangle(names):
    out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True).stdout.split("\n")
    res = {}
    for n, d in zip(names, out):
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        res[n] = d
    return res
a, b = dump(sys.argv[1]), dump(sys.argv[2])
da, db = demangle(list(a)), demangle(list(b))
ia = {da[n].replace('assemble_matrix_p1_walk<1, true>','assemble_matrix_p1_walk<1, true, false>').replace('assemble_matrix_p1_walk<2, true>','assemble_matrix_p1_walk<2, true, false>'): a[n] for n in a}
ib = {db[n]: b[n] for n in b}
same = diff = 0
for name in ia:
    if name in ib:
        if ia[name] == ib[name]: same += 1
        else:
            diff += 1; print("DIFF ", name[:150], len(ia[name]), len(ib[name]))
    else:
        print("GONE ", name[:150])
print("same", same, "diff", diff, "new", len([n for n in ib if n not in ia]))
