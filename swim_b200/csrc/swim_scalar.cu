// swim_scalar.cu — scalar Core.hs parity calls (placeholder, filled in next).
#include "swim_host.h"
