#pragma once
#include "gflags/gflags.h"
