import ctypes, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, sys, time
lib = ctypes.CDLL('/root/repo/oracle/_build/liboracle_mf.so')
ip = ctypes.POINTER(ctypes.c_int); dp = ctypes.POINTER(ctypes.c_double)
lib.oracle_mf_solve.argtypes = [ctypes.c_int, ip, ip, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, dp, dp, ctypes.c_int]
lib.oracle_mf_analyze.argtypes = [ctypes.c_int, ip, ip, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, ctypes.c_int]
def P(a, t): return a.ctypes.data_as(t)
def solve(A, b, sym_lower=0, ordering=0, matching=0, W=0, leaf=0, nref=2, verbose=0):
    A = sp.csr_matrix(A); A.sort_indices()
    rp = A.indptr.astype(np.int32); ci = A.indices.astype(np.int32); v = A.data.astype(np.float64)
    x = np.zeros(A.shape[0]); st = np.zeros(10)
    rc = lib.oracle_mf_solve(A.shape[0], P(rp,ip), P(ci,ip), P(v,dp), sym_lower, ordering, matching, W, leaf, nref, 0.0, P(b,dp), P(x,dp), P(st,dp), verbose)
    return rc, x, st
def analyze(A, sym_lower=0, ordering=0, matching=0, W=0, leaf=0, verbose=1):
    A = sp.csr_matrix(A); A.sort_indices()
    rp = A.indptr.astype(np.int32); ci = A.indices.astype(np.int32); v = A.data.astype(np.float64)
    st = np.zeros(10)
    rc = lib.oracle_mf_analyze(A.shape[0], P(rp,ip), P(ci,ip), P(v,dp), sym_lower, ordering, matching, W, leaf, P(st,dp), verbose)
    return rc, st
def lap2d(k):
    T = sp.diags([-1,2,-1],[-1,0,1],shape=(k,k))
    I = sp.identity(k)
    return (sp.kron(I,T)+sp.kron(T,I)).tocsr()
if __name__ == '__main__':
    k = int(sys.argv[1]) if len(sys.argv)>1 else 30
    A = lap2d(k); n = A.shape[0]; b = np.ones(n)
    t=time.time(); rc, x, st = solve(A, b, verbose=1); t=time.time()-t
    print('rc', rc, 'time', t, 'resid', np.linalg.norm(b-A@x)/np.linalg.norm(b), st)
    xs = spla.splu(A.tocsc()).solve(b)
    print('vs superlu', np.abs(x-xs).max()/np.abs(xs).max())
