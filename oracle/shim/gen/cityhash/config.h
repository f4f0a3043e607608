// Test-infrastructure shim (oracle/): empty cityhash config.h (third_party/cityhash_config.h.in).
#pragma once
