// swim_codec.cpp — Envelope wire codec (placeholder, filled in next).
#include "../../include/swim.h"
