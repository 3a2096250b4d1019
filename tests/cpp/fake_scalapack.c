/* TEST INFRASTRUCTURE: a stand-in for "the application's own ScaLAPACK" that sits AFTER libcosma_pxgemm.so in link order. Its pdgemm_
 * does no arithmetic: it stamps a marker into c[0] so that the test can tell which library served a call. */
void pdgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha, const double* a, const int* ia,
             const int* ja, const int* desca, const double* b, const int* ib, const int* jb, const int* descb, const double* beta, double* c,
             const int* ic, const int* jc, const int* descc) {
    (void)ta; (void)tb; (void)m; (void)n; (void)k; (void)alpha; (void)a; (void)ia; (void)ja; (void)desca; (void)b; (void)ib; (void)jb; (void)descb;
    (void)beta; (void)ic; (void)jc; (void)descc;
    c[0] = 424242.0;
}
