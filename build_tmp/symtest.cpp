#include "../juliagrid.jl_b200/csrc/symbolic.hpp"
#include <cstdio>
#include <chrono>
extern "C" int symtest(int n, const int* colptr, const int* rowidx, const int* group, const double* aval, const double* rhs, double* x, double* stats) {
    jgb::Symbolic S; jgb::SymbolicOptions opt;
    auto t0 = std::chrono::steady_clock::now();
    try { jgb::analyse(n, colptr, rowidx, group, nullptr, opt, S); } catch (std::exception& e) { printf("ERR %s\n", e.what()); return -1; }
    auto t1 = std::chrono::steady_clock::now();
    int rc = jgb::host_factor_solve(S, aval, rhs, x);
    auto t2 = std::chrono::steady_clock::now();
    stats[0]=S.nfronts; stats[1]=S.nlevels; stats[2]=S.ndepths; stats[3]=S.nnz_lu; stats[4]=S.flops; stats[5]=S.max_front; stats[6]=S.u_size; stats[7]=S.upd_size;
    stats[8]=std::chrono::duration<double>(t1-t0).count(); stats[9]=std::chrono::duration<double>(t2-t1).count();
    // histogram of front sizes
    int h[8]={0}; for (int f=0; f<S.nfronts; ++f){int nf=S.f_nf[f]; int b= nf<=2?0: nf<=4?1: nf<=8?2: nf<=16?3: nf<=32?4: nf<=64?5: nf<=128?6:7; h[b]++;}
    printf("front size hist <=2:%d <=4:%d <=8:%d <=16:%d <=32:%d <=64:%d <=128:%d >128:%d\n",h[0],h[1],h[2],h[3],h[4],h[5],h[6],h[7]);
    printf("level sizes:"); for (int l=0;l<S.nlevels;++l) printf(" %d", S.levelptr[l+1]-S.levelptr[l]); printf("\n");
    return rc;
}
