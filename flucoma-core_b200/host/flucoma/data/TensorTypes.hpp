// Aliases used throughout the reference's algorithm signatures (reference: include/flucoma/data/TensorTypes.hpp:18-31).
#pragma once
#include "FluidTensor.hpp"
#include <complex>

namespace fluid {
using RealMatrix = FluidTensor<double, 2>;
using RealMatrixView = FluidTensorView<double, 2>;
using RealVector = FluidTensor<double, 1>;
using RealVectorView = FluidTensorView<double, 1>;
using ComplexMatrix = FluidTensor<std::complex<double>, 2>;
using ComplexMatrixView = FluidTensorView<std::complex<double>, 2>;
using ComplexVector = FluidTensor<std::complex<double>, 1>;
using ComplexVectorView = FluidTensorView<std::complex<double>, 1>;
} // namespace fluid
