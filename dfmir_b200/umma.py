"""tcgen05 (UMMA) convolution engine bindings: which layer shapes run on the tensor cores and the
calls into libdfmir_b200.so for them.  Until a shape is supported here it runs on the fp32 path."""


def supported(nd, Cin, Cout, kernel, stride, pad, x, planar_out):
    return False
