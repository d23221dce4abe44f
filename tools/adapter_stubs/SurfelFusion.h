#include "stub_reference.h"
