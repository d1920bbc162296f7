import importlib, json, os, sys
sys.path.insert(0, '/root/repo')
pt = importlib.import_module("performance-test_b200")
P = pt.host.Problem("elasticity", 1, 148, 148, 149)
out = {}
for name, env in (("ring", {}), ("noring", {"PTB_ASM_RING": "0"})):
    os.environ.pop("PTB_ASM_RING", None)
    os.environ.update(env)
    c = pt.abi.Context(0)
    c.set_problem(P)
    c.assemble_matrix(); c.assemble_vector()
    c.cg_solve(kmax=60, rtol=1e-8, precond="jacobi")
    a = c.time_kernel(pt.abi.KERNEL_SPMV, 30)
    b = c.time_kernel(pt.abi.KERNEL_SPMV, 30)
    out[name] = [a, b, c.stage_ms(pt.abi.STAGE_SOLVE) / 60]
    c.close()
print(json.dumps(out))
