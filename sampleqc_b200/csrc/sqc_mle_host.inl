// host driver of the MLE variant.  STUB.
namespace {
void mle_alloc(sqc_session&, const sqc_fit_args*, sqc::MleBufs&) { fail(SQC_ERR_BAD_ARG, "MLE variant not built yet"); }
void mle_run(sqc_session*) { fail(SQC_ERR_BAD_ARG, "MLE variant not built yet"); }
void mle_fetch(sqc_session*, sqc_fit_out*) { fail(SQC_ERR_BAD_ARG, "MLE variant not built yet"); }
}
