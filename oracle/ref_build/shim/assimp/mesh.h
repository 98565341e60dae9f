#pragma once
#include "scene.h"
