"""TEST INFRASTRUCTURE. Cuts the self-contained pieces of the reference's hot path out of
/root/reference/src *where they lie* and writes them, byte for byte, into oracle/_ref/gen/*.inc
(git-ignored build output: no reference source enters the repository). oracle/ref/shim.cpp
includes the fragments and wraps them in extern "C" functions.

What is cut (anchors are checked, so a changed reference fails loudly instead of silently
extracting something else):
  mesh.cpp               num_entities / num_pdofs (:44-74) and the body of create_cube_mesh up to
                         the end of the neighbourhood search (:82-151)           -> sizing.inc
  poisson_problem.cpp    the three generic lambdas (:60-71 BC marker, :86-97 f, :100-106 g)
  elasticity_problem.cpp the two generic lambdas (:127-138 BC marker, :155-176 f)  -> lambdas.inc
  cgpoisson_problem.cpp  pack_fn / unpack_fn (:32-44)                              -> pack.inc
cg.h is not cut at all: shim.cpp includes it unchanged.

Usage: python oracle/ref/extract.py <reference_root> <out_dir>
"""
import os
import sys


def _lines(path):
    with open(path) as fh:
        return fh.read().split("\n")


def _find(lines, needle, start=0):
    for i in range(start, len(lines)):
        if needle in lines[i]:
            return i
    raise SystemExit(f"extract.py: anchor {needle!r} not found")


def _balanced(lines, i0, col0):
    """Text from (i0, col0) to the brace that closes the first '{' found after that point."""
    depth, seen, out = 0, False, []
    i, j = i0, col0
    while i < len(lines):
        line = lines[i]
        while j < len(line):
            ch = line[j]
            out.append(ch)
            if ch == "{":
                depth += 1
                seen = True
            elif ch == "}":
                depth -= 1
                if seen and depth == 0:
                    return "".join(out), i
            j += 1
        out.append("\n")
        i, j = i + 1, 0
    raise SystemExit("extract.py: unbalanced braces")


def lambdas(path, expect):
    lines = _lines(path)
    found, i = [], 0
    while True:
        try:
            i = _find(lines, "[](auto x)", i)
        except SystemExit:
            break
        text, end = _balanced(lines, i, lines[i].index("[](auto x)"))
        found.append((i + 1, end + 1, text))
        i = end + 1
    if len(found) != expect:
        raise SystemExit(f"extract.py: {path}: {len(found)} '[](auto x)' lambdas, expected {expect}")
    return found


def sizing(path):
    lines = _lines(path)
    a = _find(lines, "constexpr std::tuple<std::int64_t, std::int64_t, std::int64_t, std::int64_t>")
    b = _find(lines, "} // namespace", a)
    counts = "\n".join(lines[a:b])
    c = _find(lines, "create_cube_mesh(MPI_Comm comm", b)
    d = _find(lines, "{", c + 1)           # opening brace of the function body
    e = _find(lines, "#ifdef HAS_PARMETIS", d)
    body = "\n".join(lines[d + 1:e])
    for needle in ("const std::int64_t Nx_max = 200;", "std::size_t mindiff = 1000000;",
                   "dolfinx::MPI::size(comm)", "if (diff < mindiff)"):
        if needle not in body:
            raise SystemExit(f"extract.py: sizing body lost {needle!r}")
    return (a + 1, b, counts), (d + 2, e, body)


def pack(path):
    lines = _lines(path)
    a = _find(lines, "void pack_fn(")
    b = _find(lines, "} // namespace", a)
    return a + 1, b, "\n".join(lines[a:b])


def main(ref, out):
    src = os.path.join(ref, "src")
    os.makedirs(out, exist_ok=True)
    (ca, cb, counts), (ba, bb, body) = sizing(os.path.join(src, "mesh.cpp"))
    with open(os.path.join(out, "sizing.inc"), "w") as fh:
        fh.write(f"// generated from src/mesh.cpp:{ca}-{cb} and :{ba}-{bb}; do not commit\n")
        fh.write("namespace {\n" + counts + "\n}\n")
        fh.write("static void ref_sizing_body(MPI_Comm comm, std::size_t target_dofs, bool target_dofs_total,\n"
                 "                            std::size_t dofs_per_node, int order, std::int64_t* out)\n{\n")
        fh.write(body + "\n  out[0] = Nx; out[1] = Ny; out[2] = Nz; out[3] = r;\n}\n")
    with open(os.path.join(out, "lambdas.inc"), "w") as fh:
        names = {"poisson_problem.cpp": ("poisson_bc_marker", "poisson_f", "poisson_g"),
                 "elasticity_problem.cpp": ("elasticity_bc_marker", "elasticity_f")}
        for fname, nm in names.items():
            for (l0, l1, text), name in zip(lambdas(os.path.join(src, fname), len(nm)), nm):
                fh.write(f"// generated from src/{fname}:{l0}-{l1}; do not commit\n")
                fh.write(f"static const auto ref_{name} = {text};\n")
    a, b, text = pack(os.path.join(src, "cgpoisson_problem.cpp"))
    with open(os.path.join(out, "pack.inc"), "w") as fh:
        fh.write(f"// generated from src/cgpoisson_problem.cpp:{a}-{b}; do not commit\n")
        fh.write("namespace {\n" + text + "\n}\n")
    print("extract.py: wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
