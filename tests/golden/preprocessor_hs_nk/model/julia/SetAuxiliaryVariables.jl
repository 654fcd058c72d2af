function set_auxiliary_variables!(y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real})
@inbounds begin

end
end
