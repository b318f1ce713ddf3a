#pragma once
#include "../filesystem.hpp"
