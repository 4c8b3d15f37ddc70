/*
 * CPU oracle (C form) for the smearFEM.jl assembly path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product library (libsmearfem_b200.so) never links or calls it.
 *
 * Restates, loop for loop, the reference's algorithm (paths relative to /root/reference):
 *   gaussian_quadrature   src/fem.jl:21-31
 *   basis_function (Q1)   src/fem.jl:48-75
 *   assemble_system       src/fem.jl:135-252   element loop -> COO triplets (E,J,V), slot order `inz`
 *   sparse(E,J,V)         src/fem.jl:253       Julia stdlib SparseArrays (not in /root/reference,
 *                                              un-pinned): dims = (max E, max J); duplicates summed in
 *                                              input order; numerical zeros kept; rows ascending per column.
 *
 * PARITY STATUS: parity unpinned for K (the reference holds no golden vectors for assembly and
 * Julia is not installed); this form is cross-checked against oracle/fem_oracle.py (an independent
 * NumPy restatement) and analytic known answers in tests/.
 *
 * It doubles as the CPU baseline ("kind": "port") of bench.py: the element loop is what the
 * reference runs serially (src/fem.jl:179); `nthreads > 1` runs that loop with OpenMP over elements
 * (each element owns its COO slots, so there is no race), the COO->CSC counting sort stays serial
 * like Julia's sparse().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* src/fem.jl:21-31 (same expression order) */
void oracle_gaussian_quadrature(double a, double b, int n, double *xi, double *w) {
    if (n == 2) {
        xi[0] = -(b - a) / (2 * sqrt(3.0)) + (b + a) / 2;
        xi[1] = (b - a) / (2 * sqrt(3.0)) + (b + a) / 2;
        w[0] = (b - a) / 2;
        w[1] = (b - a) / 2;
    } else if (n == 3) {
        xi[0] = -(b - a) / (2 * sqrt(5.0 / 3.0)) + (b + a) / 2;
        xi[1] = 0.0;
        xi[2] = (b - a) / (2 * sqrt(5.0 / 3.0)) + (b + a) / 2;
        w[0] = (b - a) / 2 * 5 / 9;
        w[1] = (b - a) / 2 * 8 / 9;
        w[2] = (b - a) / 2 * 5 / 9;
    }
}

/* src/fem.jl:51-69: dN is nn x ndim, row-major here */
static void basis_q1(int ndim, double x, double e, double z, double *N, double *dN) {
    if (ndim == 3) {
        const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
        const double sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
        const double sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
        for (int a = 0; a < 8; ++a) {
            double fx = 1 + sx[a] * x, fy = 1 + sy[a] * e, fz = 1 + sz[a] * z;
            N[a] = fx * fy * fz / 8;
            dN[3 * a + 0] = sx[a] * fy * fz / 8;
            dN[3 * a + 1] = sy[a] * fx * fz / 8;
            dN[3 * a + 2] = sz[a] * fx * fy / 8;
        }
    } else {
        const double sx[4] = {-1, 1, 1, -1};
        const double sy[4] = {-1, -1, 1, 1};
        for (int a = 0; a < 4; ++a) {
            double fx = 1 + sx[a] * x, fy = 1 + sy[a] * e;
            N[a] = fx * fy / 4;
            dN[2 * a + 0] = sx[a] * fy / 4;
            dN[2 * a + 1] = sy[a] * fx / 4;
        }
    }
}

static double det_inv(int nd, const double *J, double *inv) {
    if (nd == 2) {
        double d = J[0] * J[3] - J[1] * J[2];
        inv[0] = J[3] / d; inv[1] = -J[1] / d; inv[2] = -J[2] / d; inv[3] = J[0] / d;
        return d;
    }
    double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
    double d = J[0] * c00 + J[1] * c01 + J[2] * c02;
    inv[0] = c00 / d; inv[1] = (J[2] * J[7] - J[1] * J[8]) / d; inv[2] = (J[1] * J[5] - J[2] * J[4]) / d;
    inv[3] = c01 / d; inv[4] = (J[0] * J[8] - J[2] * J[6]) / d; inv[5] = (J[2] * J[3] - J[0] * J[5]) / d;
    inv[6] = c02 / d; inv[7] = (J[1] * J[6] - J[0] * J[7]) / d; inv[8] = (J[0] * J[4] - J[1] * J[3]) / d;
    return d;
}

/*
 * src/fem.jl:135-252.  Julia layouts: NodeList ndim x nNodes column-major; IEN nEl x nn column-major,
 * 1-based; ID nNodes x nDof column-major, 1-based (ignored for nDof == 1, src/fem.jl:204-205).
 * E, J, V must hold nEl*(nn*nDof)^2 entries; V must be zero on entry (src/fem.jl:139-145).
 */
void oracle_assemble_coo_n(int64_t nEl, int ndim, int nDof, const double *NodeList, const int64_t *IEN, const int64_t *ID,
                           int64_t nNodes, double Young, double nu, int64_t *E, int64_t *J, double *V, int nthreads);

void oracle_assemble_coo(int64_t ne, int ndim, int nDof, const double *NodeList, const int64_t *IEN,
                         const int64_t *ID, int64_t nNodes, double Young, double nu, int64_t *E, int64_t *J,
                         double *V, int nthreads) {
    int64_t nEl = 1;
    for (int d = 0; d < ndim; ++d) nEl *= ne; /* src/fem.jl:179: the loop runs 1:ne^ndim */
    oracle_assemble_coo_n(nEl, ndim, nDof, NodeList, IEN, ID, nNodes, Young, nu, E, J, V, nthreads);
}

/* The same element loop over the first nEl rows of an nEl x nn connectivity (bench.py's CPU arm times a slab of element
 * layers of the full mesh as a bounded sample of the workload). */
void oracle_assemble_coo_n(int64_t nEl, int ndim, int nDof, const double *NodeList, const int64_t *IEN, const int64_t *ID,
                           int64_t nNodes, double Young, double nu, int64_t *E, int64_t *J, double *V, int nthreads) {
    const int nn = 1 << ndim, ngp = 1 << ndim, nd = nn * nDof;
    const int64_t blk = (int64_t)nd * nd;
    double xi[2], wq[2];
    oracle_gaussian_quadrature(-1, 1, 2, xi, wq);
    /* src/fem.jl:161-164, :172-176 */
    const int ix[8] = {0, 1, 1, 0, 0, 1, 1, 0}, iy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, iz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    double gx[8], gy[8], gz[8], wp[8];
    for (int g = 0; g < ngp; ++g) {
        gx[g] = xi[ix[g]]; gy[g] = xi[iy[g]]; gz[g] = xi[iz[g]];
        wp[g] = (ndim == 3) ? wq[ix[g]] * wq[iy[g]] * wq[iz[g]] : wq[ix[g]] * wq[iy[g]];
    }
    /* constitutive matrices, src/fem.jl:217 and :230 */
    double D[36];
    memset(D, 0, sizeof D);
    int ns = 0;
    if (nDof == 2) {
        ns = 3;
        D[0] = Young / (1 - nu * nu); D[1] = nu * Young / (1 - nu * nu);
        D[3] = nu * Young / (1 - nu * nu); D[4] = Young / (1 - nu * nu);
        D[8] = Young / (2 * (1 + nu));
    } else if (nDof == 3) {
        ns = 6;
        double f = Young / ((1 + nu) * (1 - 2 * nu));
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) D[6 * i + j] = (i == j ? 1 - nu : nu) * f;
        for (int i = 3; i < 6; ++i) D[6 * i + i] = (1 - 2 * nu) / 2 * f;
    }
#ifdef _OPENMP
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int64_t e = 0; e < nEl; ++e) {
        double coords[24], N[8], dN[24], Jac[9], inv[9], dNdX[24], B[6 * 24], DB[6 * 24], Ke[24 * 24];
        int64_t node[8];
        for (int a = 0; a < nn; ++a) {
            node[a] = IEN[e + a * nEl]; /* 1-based */
            for (int d = 0; d < ndim; ++d) coords[d * nn + a] = NodeList[(node[a] - 1) * ndim + d];
        }
        for (int g = 0; g < ngp; ++g) {
            basis_q1(ndim, gx[g], gy[g], gz[g], N, dN);
            /* Jac = coords*dN  (src/fem.jl:192) */
            for (int r = 0; r < ndim; ++r)
                for (int c = 0; c < ndim; ++c) {
                    double s = 0;
                    for (int a = 0; a < nn; ++a) s += coords[r * nn + a] * dN[a * ndim + c];
                    Jac[r * ndim + c] = s;
                }
            double w = wp[g] * fabs(det_inv(ndim, Jac, inv)); /* :194-195 */
            for (int a = 0; a < nn; ++a)                      /* dNdX = dN*invJ, :196 */
                for (int c = 0; c < ndim; ++c) {
                    double s = 0;
                    for (int k = 0; k < ndim; ++k) s += dN[a * ndim + k] * inv[k * ndim + c];
                    dNdX[a * ndim + c] = s;
                }
            if (nDof == 1) { /* :199-208 */
                for (int i = 0; i < nn; ++i)
                    for (int j = 0; j < nn; ++j) {
                        int64_t inz = (int64_t)nn * nn * e + nn * i + j;
                        double s = 0;
                        for (int c = 0; c < ndim; ++c) s += dNdX[i * ndim + c] * dNdX[j * ndim + c];
                        E[inz] = node[i];
                        J[inz] = node[j];
                        V[inz] += w * s;
                    }
                continue;
            }
            memset(B, 0, sizeof(double) * ns * nd);
            if (nDof == 2) { /* :211-215 */
                for (int a = 0; a < nn; ++a) {
                    B[0 * nd + 2 * a] = dNdX[2 * a];
                    B[1 * nd + 2 * a + 1] = dNdX[2 * a + 1];
                    B[2 * nd + 2 * a] = dNdX[2 * a + 1];
                    B[2 * nd + 2 * a + 1] = dNdX[2 * a];
                }
            } else { /* :219-228, Voigt xx yy zz yz xz xy */
                for (int a = 0; a < nn; ++a) {
                    double dx = dNdX[3 * a], dy = dNdX[3 * a + 1], dz = dNdX[3 * a + 2];
                    B[0 * nd + 3 * a] = dx;
                    B[1 * nd + 3 * a + 1] = dy;
                    B[2 * nd + 3 * a + 2] = dz;
                    B[3 * nd + 3 * a + 1] = dz; B[3 * nd + 3 * a + 2] = dy;
                    B[4 * nd + 3 * a] = dz;     B[4 * nd + 3 * a + 2] = dx;
                    B[5 * nd + 3 * a] = dy;     B[5 * nd + 3 * a + 1] = dx;
                }
            }
            /* Ke = B'*cMat*B*w  (:233), dense like the reference's two dgemms */
            for (int r = 0; r < ns; ++r)
                for (int c = 0; c < nd; ++c) {
                    double s = 0;
                    for (int k = 0; k < ns; ++k) s += D[r * ns + k] * B[k * nd + c];
                    DB[r * nd + c] = s;
                }
            for (int i = 0; i < nd; ++i)
                for (int j = 0; j < nd; ++j) {
                    double s = 0;
                    for (int k = 0; k < ns; ++k) s += B[k * nd + i] * DB[k * nd + j];
                    Ke[i * nd + j] = s * w;
                }
            /* scatter, :236-249; inz = |Ke|(e-1) + (iNode-1) nDof ncol + (jNode-1) nDof^2 + (iDof-1) nDof + jDof */
            for (int iN = 0; iN < nn; ++iN)
                for (int jN = 0; jN < nn; ++jN)
                    for (int iD = 0; iD < nDof; ++iD)
                        for (int jD = 0; jD < nDof; ++jD) {
                            int64_t inz = blk * e + (int64_t)iN * nDof * nd + (int64_t)jN * nDof * nDof + iD * nDof + jD;
                            E[inz] = ID[(node[iN] - 1) + (int64_t)iD * nNodes];
                            J[inz] = ID[(node[jN] - 1) + (int64_t)jD * nNodes];
                            V[inz] += Ke[(iN * nDof + iD) * nd + (jN * nDof + jD)];
                        }
        }
    }
}

/*
 * sparse(E,J,V): stable two-pass counting sort (rows, then columns) followed by an in-order fold.
 * Returns nnz; colptr must hold n+1 entries (1-based on return); rowval/nzval must hold `len`
 * entries (only the first nnz are meaningful).  m/n are returned through pm/pn.
 */
int64_t oracle_sparse(int64_t len, const int64_t *E, const int64_t *J, const double *V, int64_t *pm,
                      int64_t *pn, int64_t *colptr, int64_t *rowval, double *nzval) {
    int64_t m = 0, n = 0;
    for (int64_t t = 0; t < len; ++t) {
        if (E[t] > m) m = E[t];
        if (J[t] > n) n = J[t];
    }
    *pm = m;
    *pn = n;
    int64_t *cnt = (int64_t *)calloc((size_t)(m > n ? m : n) + 2, sizeof(int64_t));
    int64_t *p1 = (int64_t *)malloc(sizeof(int64_t) * (size_t)len);
    int64_t *p2 = (int64_t *)malloc(sizeof(int64_t) * (size_t)len);
    /* pass 1: by row */
    for (int64_t t = 0; t < len; ++t) cnt[E[t]]++;
    for (int64_t r = 1, s = 0; r <= m; ++r) { int64_t c = cnt[r]; cnt[r] = s; s += c; }
    for (int64_t t = 0; t < len; ++t) p1[cnt[E[t]]++] = t;
    /* pass 2: by column (stable) */
    memset(cnt, 0, sizeof(int64_t) * ((size_t)(m > n ? m : n) + 2));
    for (int64_t t = 0; t < len; ++t) cnt[J[t]]++;
    for (int64_t c = 1, s = 0; c <= n; ++c) { int64_t k = cnt[c]; cnt[c] = s; s += k; }
    for (int64_t t = 0; t < len; ++t) { int64_t src = p1[t]; p2[cnt[J[src]]++] = src; }
    /* fold duplicates left to right */
    int64_t nnz = 0, pc = 0, pr = 0;
    for (int64_t c = 0; c <= n; ++c) colptr[c] = 0;
    for (int64_t t = 0; t < len; ++t) {
        int64_t src = p2[t];
        if (t > 0 && J[src] == pc && E[src] == pr) {
            nzval[nnz - 1] += V[src];
        } else {
            pc = J[src]; pr = E[src];
            rowval[nnz] = pr;
            nzval[nnz] = V[src];
            colptr[pc]++; /* count in slot pc (1-based column), shifted below */
            ++nnz;
        }
    }
    /* colptr[c] currently = count of column c (c = 1..n); make it 1-based start offsets */
    int64_t s = 1;
    for (int64_t c = 1; c <= n; ++c) { int64_t k = colptr[c]; colptr[c - 1] = s; s += k; }
    colptr[n] = s;
    free(cnt); free(p1); free(p2);
    return nnz;
}

/* processors available to this process (not affected by OMP_NUM_THREADS, which launchers such as torchrun set to 1) */
int oracle_hw_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* CSR (0-based here, = the symmetric-pattern CSC shifted) SpMV y = A x, used only as the CPU solve baseline */
void oracle_spmv(int64_t n, const int64_t *colptr1, const int64_t *rowval1, const double *nzval, const double *x,
                 double *y, int nthreads) {
    /* CSC product, column-oriented like Julia's  K*x  (examples/vector3D.jl:320) */
    (void)nthreads;
    for (int64_t i = 0; i < n; ++i) y[i] = 0;
    for (int64_t c = 0; c < n; ++c) {
        double xc = x[c];
        for (int64_t p = colptr1[c] - 1; p < colptr1[c + 1] - 1; ++p) y[rowval1[p] - 1] += nzval[p] * xc;
    }
}
