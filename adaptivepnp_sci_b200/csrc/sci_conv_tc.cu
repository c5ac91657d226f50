// Dispatch of the conv entry points + (to come) the tcgen05/TMEM/TMA implicit-GEMM kernels.
#include "sci_common.cuh"

int sci_conv3x3_ref_launch(const sci_conv_desc* d, void* stream);
int sci_wgrad_ref_launch(const sci_wgrad_desc* d, void* stream);

static int check_conv_desc(const sci_conv_desc* d) {
    SCI_REQUIRE(d && d->x && d->w && d->y, "conv: null pointer");
    SCI_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv: shape");
    SCI_REQUIRE(d->Cin % 8 == 0 && d->Cout % 4 == 0, "conv: Cin % 8, Cout % 4");
    SCI_REQUIRE(d->stride == 1 || d->stride == 2, "conv: stride");
    SCI_REQUIRE(!d->pixel_shuffle || d->stride == 1, "conv: pixel_shuffle with stride 2");
    return SCI_OK;
}

extern "C" int sci_conv3x3_fwd(const sci_conv_desc* d, int impl, void* stream) {
    int rc = check_conv_desc(d);
    if (rc) return rc;
    if (impl == SCI_CONV_REF) return sci_conv3x3_ref_launch(d, stream);
    return sci_fail(SCI_EUNSUPPORTED, "conv: tensor-core path not built yet");
}

extern "C" int sci_conv3x3_dgrad(const sci_conv_desc* d, int impl, void* stream) {
    return sci_conv3x3_fwd(d, impl, stream);
}

extern "C" int sci_conv3x3_wgrad(const sci_wgrad_desc* d, int impl, void* stream) {
    SCI_REQUIRE(d && d->x && d->dz && d->dw, "wgrad: null pointer");
    SCI_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && (d->stride == 1 || d->stride == 2),
                "wgrad: shape");
    SCI_REQUIRE(d->Cin % 4 == 0 && d->Cout % 4 == 0, "wgrad: channels % 4");
    if (impl == SCI_CONV_REF) return sci_wgrad_ref_launch(d, stream);
    return sci_fail(SCI_EUNSUPPORTED, "wgrad: tensor-core path not built yet");
}
