// TEST INFRASTRUCTURE. Empty stand-in for <dolfinx/common/IndexMap.h>: /root/reference/src/cg.h
// includes it but uses nothing from it. See oracle/ref/README.md.
#pragma once
