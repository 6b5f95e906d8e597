// Type-only stand-in for the Bullet classes named in cosa0481/DRRT's headers
// (TEST INFRASTRUCTURE, see oracle/Makefile target `ref`).  Bullet is not in
// this image.  These let the reference's translation units compile; nothing on
// the path oracle/_ref checks (KD tree, ghost points, Dubins trajectory, the
// analytic ExplicitEdgeCheck2D) calls into them -- the collision world reports
// no contacts, and the driver never routes a verdict through it.
#ifndef DRRT_BULLET_STUB_H
#define DRRT_BULLET_STUB_H
typedef double btScalar;
struct btVector3 { btScalar v[3]; btVector3() {} btVector3(btScalar a, btScalar b, btScalar c) { v[0]=a; v[1]=b; v[2]=c; } };
struct btQuaternion { btQuaternion() {} btQuaternion(btScalar, btScalar, btScalar) {} };
struct btTransform { void setOrigin(const btVector3 &) {} void setRotation(const btQuaternion &) {} };
struct btCollisionShape { virtual ~btCollisionShape() {} };
struct btConvexHullShape : btCollisionShape { void addPoint(const btVector3 &) {} };
struct btCylinderShape : btCollisionShape { btCylinderShape(const btVector3 &) {} };
struct btCollisionObject {
  btTransform t;
  void setCollisionShape(btCollisionShape *) {}
  btTransform &getWorldTransform() { return t; }
  const btTransform &getWorldTransform() const { return t; }
};
struct btManifoldPoint { btScalar getDistance() const { return 1; } };
struct btPersistentManifold {
  const btCollisionObject *getBody0() const { return 0; }
  const btCollisionObject *getBody1() const { return 0; }
  void refreshContactPoints(const btTransform &, const btTransform &) {}
  int getNumContacts() const { return 0; }
  btManifoldPoint &getContactPoint(int) { static btManifoldPoint p; return p; }
};
struct btCollisionConfiguration { virtual ~btCollisionConfiguration() {} };
struct btDefaultCollisionConfiguration : btCollisionConfiguration {};
struct btDispatcher {
  int getNumManifolds() const { return 0; }
  btPersistentManifold *getManifoldByIndexInternal(int) { static btPersistentManifold m; return &m; }
};
struct btCollisionDispatcher : btDispatcher { btCollisionDispatcher(btCollisionConfiguration *) {} };
struct btBroadphaseInterface { virtual ~btBroadphaseInterface() {} };
struct bt32BitAxisSweep3 : btBroadphaseInterface {
  bt32BitAxisSweep3(const btVector3 &, const btVector3 &, int = 0, void * = 0, bool = false) {}
};
struct btCollisionWorld {
  btDispatcher *d;
  btCollisionWorld(btDispatcher *dd, btBroadphaseInterface *, btCollisionConfiguration *) : d(dd) {}
  void addCollisionObject(btCollisionObject *) {}
  void removeCollisionObject(btCollisionObject *) {}
  void performDiscreteCollisionDetection() {}
  btDispatcher *getDispatcher() { return d; }
};
#endif
