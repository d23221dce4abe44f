// Stand-in header (TEST INFRASTRUCTURE): see ../cvshim.hpp / ../../cvshim.hpp
#pragma once
#include "cvshim.hpp"
