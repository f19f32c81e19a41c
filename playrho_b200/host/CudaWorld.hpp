// CudaWorld.hpp -- the B200-native world type that plugs into playrho::d2::World's type-erased
// interface next to AabbTreeWorld:   playrho::d2::World world{playrho::b200::CudaWorld{conf}};
//
// World{T} (Library/include/playrho/d2/World.hpp:739-745) wraps any copy-constructible,
// equality-comparable T in detail::WorldModel<T> (detail/WorldModel.hpp:36-370), which forwards
// every virtual of detail::WorldConcept (detail/WorldConcept.hpp:66-408) to an UNQUALIFIED free
// function found by ADL on T. This header declares T and those 52 free functions with the same
// names, argument meaning and exception behaviour as AabbTreeWorld's (d2/AabbTreeWorld.hpp:120-534).
//
// CudaWorld is a thin C++17 shim: entity VALUES (Body, Shape) are mirrored on the host because the
// L4 helpers do read-modify-write of whole values; all per-step state lives in device
// SoA buffers behind the C ABI of include/playrho_b200.h, and this file calls ONLY that C ABI.
// There is no CPU fallback: constructing a CudaWorld without a CUDA device throws.
#ifndef PLAYRHO_B200_CUDAWORLD_HPP
#define PLAYRHO_B200_CUDAWORLD_HPP

#include <playrho/d2/World.hpp>
#include <playrho/d2/WorldConf.hpp>
#include <playrho/d2/DynamicTree.hpp>
#include <playrho/d2/Body.hpp>
#include <playrho/d2/Shape.hpp>
#include <playrho/d2/Joint.hpp>
#include <playrho/d2/Manifold.hpp>
#include <playrho/Contact.hpp>
#include <playrho/StepConf.hpp>
#include <playrho/StepStats.hpp>

#include <optional>
#include <tuple>
#include <vector>

struct pr_world; // include/playrho_b200.h

namespace playrho::b200 {

using namespace playrho::d2;

class CudaWorld {
public:
    /// TOI event scheduling, see pr_set_toi_mode in include/playrho_b200.h.
    enum class ToiMode : unsigned { SerialExact = 0, Parallel = 1 };

    explicit CudaWorld(const WorldConf& conf = WorldConf{});
    CudaWorld(const CudaWorld& other);            // deep copy: device-to-device arena copy (pr_clone)
    CudaWorld(CudaWorld&& other) noexcept;
    CudaWorld& operator=(const CudaWorld& other);
    CudaWorld& operator=(CudaWorld&& other) noexcept;
    ~CudaWorld() noexcept;

    void SetToiMode(ToiMode mode);
    /// Listener calls lost because a step produced more contact events than the device event buffers held (the step itself is
    /// applied, the events that were kept are delivered, and the buffers are grown for the following steps).
    std::uint64_t GetDroppedEventCount() const noexcept;

    // ---- implementation detail (public for the free functions below; not part of the World API)
    struct Impl;
    Impl* m = nullptr;
};

bool operator==(const CudaWorld& lhs, const CudaWorld& rhs);
bool operator!=(const CudaWorld& lhs, const CudaWorld& rhs);

// Listeners: begin/end/post-solve calls are recorded on the device during Step (pr_events) and made right after it, in
// (kind, contact id) order; pre-solve is stored but never invoked (it needs a synchronous callback inside the step).
void SetShapeDestructionListener(CudaWorld& world, ShapeFunction listener) noexcept;
void SetDetachListener(CudaWorld& world, BodyShapeFunction listener) noexcept;
void SetJointDestructionListener(CudaWorld& world, JointFunction listener) noexcept;
void SetBeginContactListener(CudaWorld& world, ContactFunction listener) noexcept;
void SetEndContactListener(CudaWorld& world, ContactFunction listener) noexcept;
void SetPreSolveContactListener(CudaWorld& world, ContactManifoldFunction listener) noexcept;
void SetPostSolveContactListener(CudaWorld& world, ContactImpulsesFunction listener) noexcept;

// Miscellaneous
std::optional<pmr::StatsResource::Stats> GetResourceStats(const CudaWorld& world) noexcept;
void Clear(CudaWorld& world) noexcept;
StepStats Step(CudaWorld& world, const StepConf& conf);
bool IsStepComplete(const CudaWorld& world) noexcept;
bool GetSubStepping(const CudaWorld& world) noexcept;
void SetSubStepping(CudaWorld& world, bool flag) noexcept;
const DynamicTree& GetTree(const CudaWorld& world) noexcept;
bool IsLocked(const CudaWorld& world) noexcept;
void ShiftOrigin(CudaWorld& world, const Length2& newOrigin);
Interval<Positive<Length>> GetVertexRadiusInterval(const CudaWorld& world) noexcept;
Frequency GetInvDeltaTime(const CudaWorld& world) noexcept;

// Bodies
BodyCounter GetBodyRange(const CudaWorld& world) noexcept;
std::vector<BodyID> GetBodies(const CudaWorld& world);
std::vector<BodyID> GetBodiesForProxies(const CudaWorld& world);
BodyID CreateBody(CudaWorld& world, Body body = Body{});
Body GetBody(const CudaWorld& world, BodyID id);
void SetBody(CudaWorld& world, BodyID id, Body value);
void Destroy(CudaWorld& world, BodyID id);
std::vector<std::pair<BodyID, JointID>> GetJoints(const CudaWorld& world, BodyID id);
std::vector<std::tuple<ContactKey, ContactID>> GetContacts(const CudaWorld& world, BodyID id);
std::vector<ShapeID> GetShapes(const CudaWorld& world, BodyID id);

// Joints
JointCounter GetJointRange(const CudaWorld& world) noexcept;
std::vector<JointID> GetJoints(const CudaWorld& world);
JointID CreateJoint(CudaWorld& world, Joint def);
Joint GetJoint(const CudaWorld& world, JointID id);
void SetJoint(CudaWorld& world, JointID id, Joint def);
void Destroy(CudaWorld& world, JointID id);

// Shapes
ShapeCounter GetShapeRange(const CudaWorld& world) noexcept;
ShapeID CreateShape(CudaWorld& world, Shape def);
Shape GetShape(const CudaWorld& world, ShapeID id);
void SetShape(CudaWorld& world, ShapeID id, Shape def);
void Destroy(CudaWorld& world, ShapeID id);

// Contacts
ContactCounter GetContactRange(const CudaWorld& world) noexcept;
std::vector<KeyedContactID> GetContacts(const CudaWorld& world);
Contact GetContact(const CudaWorld& world, ContactID id);
void SetContact(CudaWorld& world, ContactID id, Contact value);
Manifold GetManifold(const CudaWorld& world, ContactID id);
void SetManifold(CudaWorld& world, ContactID id, const Manifold& value);

} // namespace playrho::b200

#endif
