#include <bullet/btBulletCollisionCommon.h>
