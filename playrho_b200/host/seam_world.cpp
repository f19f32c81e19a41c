// seam_world.cpp -- the link seam: World::World(const WorldConf&) is the ONLY place the
// reference names AabbTreeWorld, and it is defined out of line in Library/source/playrho/d2/World.cpp:28.
// Linking this translation unit INSTEAD of World.cpp.o makes every unmodified reference program that
// constructs World{} / World{WorldConf} (HelloWorld, UnitTests, Benchmark, Testbed) run on the
// device-resident world. The four other members World.cpp defines are restated below (they do not depend
// on the payload type): World.cpp:30-57.
#include "CudaWorld.hpp"
#include <playrho/d2/World.hpp>
#include <playrho/d2/WorldBody.hpp>

namespace playrho::d2 {

World::World(const WorldConf& def) : World{::playrho::b200::CudaWorld{def}} {}

World::World(const World& other) : m_impl{other.m_impl ? other.m_impl->Clone_() : nullptr} {}

World::~World() noexcept
{
    if (m_impl) m_impl->Clear_(); // destruction listeners run while the world is still valid
}

World& World::operator=(const World& other)
{
    m_impl = other.m_impl ? other.m_impl->Clone_() : nullptr;
    return *this;
}

BodyID CreateBody(World& world, const Body& body, bool resetMassData)
{
    const auto id = world.m_impl->CreateBody_(body);
    if (resetMassData) ResetMassData(world, id);
    return id;
}

} // namespace playrho::d2
