// Build shim (ours, not reference code): the reference's CudaMemory.h includes <GetSet/StringType.hxx>
// from the external LibGetSet, which is not vendored. Its vendored copy of the same header lives under
// code/HeaderOnly/NRRD/ in the reference tree; forward to it where it lies.
#include <NRRD/StringType.hxx>
