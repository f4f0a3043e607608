// Test-infrastructure shim (oracle/): what CMake's configure_file would generate from
// src/db/defs.h.in for a non-release, non-coverage build. GIT_SHA1 salts the JIT cache key
// (src/codegen/compiler.cc:98-100).
#pragma once
#define VIYA_VERSION "1.0.0-beta"
#define VIYA_IS_RELEASE 0
#define CODE_COVERAGE 0
#define GIT_SHA1 "56d9b9836a57a36483bee98e6bc79f79e2f5c772"
