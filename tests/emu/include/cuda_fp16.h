// empty: nccl.h includes it; the emulator build needs none of it (TEST INFRASTRUCTURE)
#pragma once
