// What cmake's configure_file would generate from the reference's config.in.h:10-12 (CMakeLists.txt:9-13,42).
#ifndef SARW_CONFIG_H
#define SARW_CONFIG_H
#include <core/Util.h>
#define PACKAGE_NAME "sarw"
#define VERSION_STRING "0.2.alpha"
#define GIT_HASH "standin"
#endif
