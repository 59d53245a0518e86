#pragma once
#include <map>
#include <string>
#include <vector>
