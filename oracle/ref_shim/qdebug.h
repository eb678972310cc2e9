#pragma once
#include <QtCore/QtGlobal>
