/* ref_select.cpp -- backend switch of libasif_ref_b200.so (TEST INFRASTRUCTURE ONLY; see qp_select_shim.h). */
extern "C" {
int g_ref_qp_backend = 0;
/* 0 = OSQP stand-in (as libasif_ref.so), 1 = ASIF::QPWrapperB200; applies to filters created afterwards */
void ref_select_backend(int backend) { g_ref_qp_backend = backend; }
int ref_has_b200_backend(void) { return 1; }
}
