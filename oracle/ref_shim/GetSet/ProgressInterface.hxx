// Build shim (ours): forwards to the reference's vendored copy, see StringType.hxx in this directory.
#include <Utils/ProgressInterface.hxx>
